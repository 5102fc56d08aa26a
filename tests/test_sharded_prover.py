"""The sharded prover (csrc/host/prover_sharded.hpp) must produce the very same proof as the single-rank prover.
CPU: the driver on 1, 2, 4 and 8 in-process ranks (threads) over the oracle backend vs the golden proofs.
GPU: the driver at world 1 through the CUDA backend (row-range kernels, views, gathers) vs the golden proofs; the multi-rank
NCCL path is exercised by tools/run_sharded_prove.py under torchrun (see profiles/)."""
import ctypes
import hashlib
import json
import os

import pytest

import proof_canon

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "proof_hashes.json")))


def source(name, g):
    return g["code"].encode() if g["code"] else open(os.path.join(ROOT, "tests", "golden", "programs", name + ".bf"), "rb").read()


@pytest.mark.parametrize("world", [1, 2, 4])
@pytest.mark.parametrize("name", ["with_input", "jump_mid", "a-bc"])
def test_sharded_driver_on_thread_ranks(orc, name, world):
    g = GOLD[name]
    lib = orc.lib
    lib.orc_prove_sharded_json.restype = ctypes.c_void_p
    lib.orc_last_error.restype = ctypes.c_char_p
    stdin = bytes.fromhex(g["stdin_hex"])
    p = lib.orc_prove_sharded_json(source(name, g), stdin, ctypes.c_size_t(len(stdin)), ctypes.c_uint32(g["log_max_rows"]), world, 1)
    assert p, lib.orc_last_error()
    js = ctypes.string_at(p)
    lib.orc_free(ctypes.c_void_p(p))
    proof_canon.check(js, g)


def test_sharded_driver_eight_thread_ranks(orc):
    """World 8 (w = 3): more ranks than tables of some sizes, 13 tables dealt over 8 ranks, replicated FRI tail from log 6."""
    g = GOLD["a-bc"]
    lib = orc.lib
    lib.orc_prove_sharded_json.restype = ctypes.c_void_p
    lib.orc_last_error.restype = ctypes.c_char_p
    stdin = bytes.fromhex(g["stdin_hex"])
    p = lib.orc_prove_sharded_json(source("a-bc", g), stdin, ctypes.c_size_t(len(stdin)), ctypes.c_uint32(g["log_max_rows"]), 8, 1)
    assert p, lib.orc_last_error()
    js = ctypes.string_at(p)
    lib.orc_free(ctypes.c_void_p(p))
    proof_canon.check(js, g)


def test_sharded_driver_hello_kakarot_two_ranks(orc):
    g = GOLD["hello_kakarot"]
    lib = orc.lib
    lib.orc_prove_sharded_json.restype = ctypes.c_void_p
    p = lib.orc_prove_sharded_json(source("hello_kakarot", g), b"", ctypes.c_size_t(0), ctypes.c_uint32(g["log_max_rows"]), 2, 1)
    assert p
    js = ctypes.string_at(p)
    lib.orc_free(ctypes.c_void_p(p))
    proof_canon.check(js, g)


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(GOLD))
def test_cuda_sharded_driver_world1(pkg, be, name):
    g = GOLD[name]
    proof = pkg.prove_brainfuck_sharded(be, None, source(name, g), bytes.fromhex(g["stdin_hex"]), g["log_max_rows"])
    proof.verify()
    js = proof.json().encode()
    proof_canon.check(js, g)
