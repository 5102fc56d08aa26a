"""Golden proof vectors (tests/golden/proof_hashes.json, made by tests/golden/make_proof_hashes.py with the CPU oracle).
CPU: the oracle still reproduces them (small cases).  GPU: the CUDA prover reproduces every one of them."""
import ctypes
import hashlib
import json
import os

import pytest

import proof_canon

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "proof_hashes.json")))


def source(name, g):
    return g["code"].encode() if g["code"] else open(os.path.join(ROOT, "tests", "golden", "programs", name.split("@")[0] + ".bf"), "rb").read()


@pytest.mark.parametrize("name", ["with_input", "no_input", "jump_mid", "a-bc", "hello_kakarot"])
def test_oracle_reproduces_golden_proof(orc, name):
    g = GOLD[name]
    lib = orc.lib
    lib.orc_prove_json.restype = ctypes.c_void_p
    stdin = bytes.fromhex(g["stdin_hex"])
    p = lib.orc_prove_json(source(name, g), stdin, ctypes.c_size_t(len(stdin)), ctypes.c_uint32(g["log_max_rows"]), 1)
    assert p
    js = ctypes.string_at(p)
    lib.orc_free(ctypes.c_void_p(p))
    proof_canon.check(js, g)


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(GOLD))
def test_cuda_prover_reproduces_golden_proof(pkg, be, name):
    g = GOLD[name]
    proof = pkg.prove_brainfuck(be, source(name, g), bytes.fromhex(g["stdin_hex"]), g["log_max_rows"])
    proof.verify()
    js = proof.json().encode()
    proof_canon.check(js, g)
    # the device path with the host tables built first gives the same proof
    js2 = pkg.prove_brainfuck(be, source(name, g), bytes.fromhex(g["stdin_hex"]), g["log_max_rows"], overlap_host=False).json().encode()
    assert js2 == js


def test_proof_json_has_the_shape_serde_gives_the_reference_types(orc):
    """`brainfuck_prover verify` does `serde_json::from_str::<BrainfuckProof<_>>` (bin/brainfuck_prover.rs:145-152).  What the
    reference's own type definitions fix about that text: the three top-level fields (brainfuck_air/mod.rs:71-76), the 13
    component names in declaration order (:78-93, :170-184), `Claim { log_size, _marker }` with the PhantomData serialised
    as null and REQUIRED on the way back (components/mod.rs:85-93), `InteractionClaim { claimed_sum }` (:70-76)."""
    import json
    lib = orc.lib
    lib.orc_prove_json.restype = ctypes.c_void_p
    p = lib.orc_prove_json(b"+>,<[>+.<-]", b"\x01", ctypes.c_size_t(1), ctypes.c_uint32(10), 0)
    raw = ctypes.string_at(p)
    lib.orc_free(ctypes.c_void_p(p))
    assert b" " not in raw and b"\n" not in raw                      # serde_json::to_string is compact
    js = json.loads(raw)
    names = ["memory", "instruction", "program", "processor", "jump_if_not_zero", "jump_if_zero", "input_instruction",
             "left_instruction", "minus_instruction", "output_instruction", "plus_instruction", "right_instruction", "end_of_execution"]
    assert list(js) == ["claim", "interaction_claim", "proof"]
    assert list(js["claim"]) == names and list(js["interaction_claim"]) == names
    for n in names:
        assert list(js["claim"][n]) == ["log_size", "_marker"] and js["claim"][n]["_marker"] is None
        cs = js["interaction_claim"][n]["claimed_sum"]
        assert list(js["interaction_claim"][n]) == ["claimed_sum"] and [len(cs), len(cs[0]), len(cs[1])] == [2, 2, 2]
    s = js["proof"]
    assert list(s) == ["commitments", "sampled_values", "decommitments", "queried_values", "proof_of_work", "fri_proof"]
    assert len(s["commitments"]) == 4 and all(len(h) == 32 and all(0 <= b < 256 for b in h) for h in s["commitments"])
    assert list(s["fri_proof"]) == ["first_layer", "inner_layers", "last_layer_poly"]
    assert list(s["fri_proof"]["first_layer"]) == ["fri_witness", "decommitment", "commitment"]
    assert list(s["decommitments"][0]) == ["hash_witness", "column_witness"]
    assert list(s["fri_proof"]["last_layer_poly"]) == ["coeffs", "log_size"] and s["fri_proof"]["last_layer_poly"]["log_size"] == 0


HELLO_WORLD = (b"++++++++++[>+++++++>++++++++++>+++>+<<<<-]>++.>+.+++++++..+++.>++.<<+++++++++++++++.>.+++.------.--------.>+.>.")


@pytest.mark.parametrize("code,stdin,out,lmr", [
    (b"+++>,<[>+.<-]", b"\x01", bytes([2, 3, 4]), 14),            # test_proof
    (HELLO_WORLD, b"", b"Hello World!\n", 16),                      # test_proof_hello_world
    (b"+++><[>+<-]", b"", b"", 14),                                 # test_proof_no_input
    (b"++[-]+.", b"", bytes([1]), 14),                              # test_proof_jump_middle_of_program
])
def test_reference_end_to_end_programs(orc, code, stdin, out, lmr):
    """The four prove -> verify round trips of crates/brainfuck_prover/src/brainfuck_air/mod.rs:804-858, program text and
    input exactly as there (LOG_MAX_ROWS smaller than the reference's test value of 20 to keep the scalar oracle quick)."""
    import host_model as H
    regs, got = H.execute(H.compile_bf(code), stdin, max_steps=20000)
    assert got == out
    lib = orc.lib
    lib.orc_prove_json.restype = ctypes.c_void_p
    lib.orc_last_error.restype = ctypes.c_char_p
    p = lib.orc_prove_json(code, stdin, ctypes.c_size_t(len(stdin)), ctypes.c_uint32(lmr), 1)      # 1: also verify
    assert p, lib.orc_last_error()
    lib.orc_free(ctypes.c_void_p(p))


def test_proof_json_round_trips_through_the_parser(orc, pkg):
    """sbf_proof_from_json o sbf_proof_json is the identity, and sbf_verify_json accepts the oracle's proof text with the right
    LOG_MAX_ROWS and rejects it with another one or when it is malformed (host code: no GPU needed)."""
    lib = pkg.load_library()
    olib = orc.lib
    olib.orc_prove_json.restype = ctypes.c_void_p
    p = olib.orc_prove_json(b"+>,<[>+.<-]", b"\x01", ctypes.c_size_t(1), ctypes.c_uint32(10), 0)
    raw = ctypes.string_at(p)
    olib.orc_free(ctypes.c_void_p(p))
    assert lib.sbf_verify_json(raw, ctypes.c_uint32(10)) == 0
    assert lib.sbf_verify_json(raw, ctypes.c_uint32(11)) != 0
    pr = pkg.Proof.from_json(lib, raw.decode(), 10)
    assert pr.json().encode() == raw
    pr.verify()
    for bad in (raw[:-1], raw.replace(b'"claim"', b'"clam"', 1), raw.replace(b'"_marker":null', b'"_marker":0', 1), raw + b" x",
                raw.replace(b'"proof_of_work":', b'"proof_of_work":-', 1)):
        assert lib.sbf_verify_json(bad, ctypes.c_uint32(10)) != 0
