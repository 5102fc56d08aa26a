"""Golden proof vectors (tests/golden/proof_hashes.json, made by tests/golden/make_proof_hashes.py with the CPU oracle).
CPU: the oracle still reproduces them (small cases).  GPU: the CUDA prover reproduces every one of them."""
import ctypes
import hashlib
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "proof_hashes.json")))


def source(name, g):
    return g["code"].encode() if g["code"] else open(os.path.join(ROOT, "tests", "golden", "programs", name + ".bf"), "rb").read()


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
    assert len(js) == g["proof_bytes"] and hashlib.sha256(js).hexdigest() == g["sha256"]


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(GOLD))
def test_cuda_prover_reproduces_golden_proof(pkg, be, name):
    g = GOLD[name]
    proof = pkg.prove_brainfuck(be, source(name, g), bytes.fromhex(g["stdin_hex"]), g["log_max_rows"])
    proof.verify()
    js = proof.json().encode()
    assert len(js) == g["proof_bytes"] and hashlib.sha256(js).hexdigest() == g["sha256"]
    # the device path with the host tables built first gives the same proof
    js2 = pkg.prove_brainfuck(be, source(name, g), bytes.fromhex(g["stdin_hex"]), g["log_max_rows"], overlap_host=False).json().encode()
    assert js2 == js
