"""The independent pure-Python verifier (tests/py_verifier.py — hashlib + integers, no code shared with csrc/host/) must accept
the proofs of both provers and reject tampered ones.  This is the test that fails if prover.hpp and verifier.hpp drift
together: transcript order, mask points, column order, FRI fold positions, witness order and the JSON shape are all restated
there from the protocol description.  (Parity with upstream Stwo itself stays unpinned: no Rust toolchain anywhere here.)"""
import ctypes
import json
import os

import pytest

import py_verifier
from test_host_tables import load

GOLD = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "proof_hashes.json")))


def oracle_proof(orc, code, stdin, lmr):
    lib = orc.lib
    lib.orc_prove_json.restype = ctypes.c_void_p
    lib.orc_last_error.restype = ctypes.c_char_p
    p = lib.orc_prove_json(code, stdin, ctypes.c_size_t(len(stdin)), ctypes.c_uint32(lmr), 0)
    assert p, lib.orc_last_error()
    js = ctypes.string_at(p)
    lib.orc_free(ctypes.c_void_p(p))
    return js


def case(name):
    g = GOLD[name]
    code = g["code"].encode() if g["code"] else load(name.split("@")[0] + ".bf")
    return code, bytes.fromhex(g["stdin_hex"]), g["log_max_rows"]


SMALL = ["with_input", "no_input", "jump_mid", "a-bc"]


@pytest.mark.parametrize("name", SMALL)
def test_oracle_proofs_are_accepted(orc, name):
    code, stdin, lmr = case(name)
    assert py_verifier.verify(oracle_proof(orc, code, stdin, lmr), lmr)


def tampered(js):
    """one corrupted field per case, mirroring sbf_proof_tamper's eight cases plus structure-level ones"""
    def edit(fn):
        p = json.loads(js)
        fn(p)
        return json.dumps(p, separators=(",", ":"))
    s = lambda p: p["proof"]
    yield "claimed_sum", edit(lambda p: p["interaction_claim"]["memory"]["claimed_sum"][0].__setitem__(0, p["interaction_claim"]["memory"]["claimed_sum"][0][0] ^ 1))
    yield "sampled value", edit(lambda p: s(p)["sampled_values"][1][0][0][0].__setitem__(0, s(p)["sampled_values"][1][0][0][0][0] ^ 1))
    yield "queried value", edit(lambda p: s(p)["queried_values"][1][0].__setitem__(0, s(p)["queried_values"][1][0][0] ^ 1))
    yield "fri witness", edit(lambda p: s(p)["fri_proof"]["first_layer"]["fri_witness"][0][0].__setitem__(0, s(p)["fri_proof"]["first_layer"]["fri_witness"][0][0][0] ^ 1))
    yield "proof of work", edit(lambda p: s(p).__setitem__("proof_of_work", s(p)["proof_of_work"] + 1))
    yield "hash witness", edit(lambda p: s(p)["decommitments"][1]["hash_witness"][0].__setitem__(0, s(p)["decommitments"][1]["hash_witness"][0][0] ^ 1))
    yield "last layer", edit(lambda p: s(p)["fri_proof"]["last_layer_poly"]["coeffs"][0][0].__setitem__(0, s(p)["fri_proof"]["last_layer_poly"]["coeffs"][0][0][0] ^ 1))
    yield "commitment", edit(lambda p: s(p)["commitments"][2].__setitem__(0, s(p)["commitments"][2][0] ^ 1))
    yield "inner layer witness", edit(lambda p: s(p)["fri_proof"]["inner_layers"][1]["fri_witness"][0][1].__setitem__(1, s(p)["fri_proof"]["inner_layers"][1]["fri_witness"][0][1][1] ^ 1))
    yield "log size", edit(lambda p: p["claim"]["program"].__setitem__("log_size", p["claim"]["program"]["log_size"] + 1))
    yield "column witness", edit(lambda p: s(p)["decommitments"][0]["column_witness"].__setitem__(0, s(p)["decommitments"][0]["column_witness"][0] ^ 1)
                                 if s(p)["decommitments"][0]["column_witness"] else s(p)["decommitments"][0]["hash_witness"].pop())
    yield "swapped sampled columns", edit(lambda p: s(p)["sampled_values"][1].__setitem__(slice(0, 2), s(p)["sampled_values"][1][1::-1]))


def test_tampered_proofs_are_rejected(orc):
    code, stdin, lmr = case("with_input")
    js = oracle_proof(orc, code, stdin, lmr)
    assert py_verifier.verify(js, lmr)
    n = 0
    for what, bad in tampered(js):
        with pytest.raises(py_verifier.Reject):
            py_verifier.verify(bad, lmr)
            pytest.fail(f"accepted a proof with a corrupted {what}")
        n += 1
    assert n == 12


def test_a_proof_made_for_another_log_max_rows_is_rejected(orc):
    code, stdin, lmr = case("no_input")
    js = oracle_proof(orc, code, stdin, lmr)
    with pytest.raises(py_verifier.Reject):
        py_verifier.verify(js, lmr + 1)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["hello_kakarot", "collatz"])
def test_cuda_proofs_are_accepted_by_the_independent_verifier(pkg, be, name):
    code, stdin, lmr = case(name)
    pr = pkg.prove_brainfuck(be, code, stdin, lmr)
    assert py_verifier.verify(pr.json(), lmr)
    for what in range(8):   # the library's own tamper hook: every case must be rejected here as well
        bad = pr.tamper(what)
        with pytest.raises(py_verifier.Reject):
            py_verifier.verify(bad.json(), lmr)
