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
    return g["code"].encode() if g["code"] else open(os.path.join(ROOT, "tests", "golden", "programs", name.split("@")[0] + ".bf"), "rb").read()


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
    proof = pkg.prove_brainfuck_sharded(be, None, source(name, g), bytes.fromhex(g["stdin_hex"]), g["log_max_rows"],
                                        overlap_host=(len(name) % 2 == 0))   # both orders of VM / preprocessed phase
    proof.verify()
    js = proof.json().encode()
    proof_canon.check(js, g)


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 4])
def test_cuda_sharded_driver_over_nccl(world):
    """The real multi-rank path: `world` processes, one GPU each, under torch.distributed.run over NCCL (column->row exchange per
    tree — NCCL all-to-all in the first fib19 proof, peer stores into the IPC-mapped receive windows in the second —
    all-gathered sub-roots, the device-side FRI transcript and tail).  Every rank must reproduce the golden proofs of with_input, a-bc,
    hello_kakarot and collatz, and the ranks must agree on fib19.  Skipped on a box with fewer GPUs."""
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs, this box has {torch.cuda.device_count()}")
    port = 29600 + world
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tools", "run_sharded_prove.py"), "--fib19-proofs", "2"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    lines = [json.loads(l) for l in r.stdout.splitlines() if l.startswith("{")]
    small = [l for l in lines if l["program"] != "fib19"]
    assert len(small) == 4 and all(l["matches_golden"] and l["world"] == world for l in small)
    fib = [l for l in lines if l["program"] == "fib19"]
    assert len(fib) == 2 and all(l["ranks_agree"] for l in fib)
    assert fib[0]["sha256"] == fib[1]["sha256"]
    assert fib[0]["sha256"] == json.load(open(os.path.join(ROOT, "tests", "golden", "fib19_wire_sha256.json")))["sha256"]


@pytest.mark.parametrize("min_log", [7, 9, 30])
def test_sharded_driver_with_small_columns_replicated(orc, min_log, monkeypatch):
    """ProverConfig::shard_min_log: columns below 2^min_log rows are replicated and their trees hashed whole on every rank
    (what the CUDA path does below 2^16 rows); 30 replicates everything.  The proof must not change."""
    monkeypatch.setenv("ORC_SHARD_MIN_LOG", str(min_log))
    g = GOLD["a-bc"]
    lib = orc.lib
    lib.orc_prove_sharded_json.restype = ctypes.c_void_p
    lib.orc_last_error.restype = ctypes.c_char_p
    stdin = bytes.fromhex(g["stdin_hex"])
    p = lib.orc_prove_sharded_json(source("a-bc", g), stdin, ctypes.c_size_t(len(stdin)), ctypes.c_uint32(g["log_max_rows"]), 4, 1)
    assert p, lib.orc_last_error()
    js = ctypes.string_at(p)
    lib.orc_free(ctypes.c_void_p(p))
    proof_canon.check(js, g)


def test_columns_are_dealt_longest_first(orc):
    """assign_owners (prover_sharded.hpp): columns in descending size, each to the least-loaded rank.  The interaction tree of
    fib19 — 4 coordinate columns of 2^24 words, 16 of 2^22, 16 of 2^20, 8 of 2^19, 8 of 2^11, 12 of 2^4 — must come out balanced
    on 2, 4 and 8 ranks (a round-robin over the sorted list gives 6 : 3 on eight ranks), every rank must compute the same
    assignment, and one rank owns everything at world 1."""
    import numpy as np
    lib = orc.lib
    logs = np.array([24] * 4 + [22] * 16 + [11] * 4 + [19] * 8 + [4] * 12 + [20] * 16 + [11] * 4, dtype=np.uint32)
    for world in (1, 2, 4, 8):
        out = np.zeros(len(logs), dtype=np.int32)
        lib.orc_assign_owners(logs.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32)), ctypes.c_size_t(len(logs)), world,
                              out.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)))
        again = np.zeros_like(out)
        lib.orc_assign_owners(logs.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32)), ctypes.c_size_t(len(logs)), world,
                              again.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)))
        assert (out == again).all() and out.min() >= 0 and out.max() < world
        load = np.zeros(world)
        for lg, o in zip(logs, out):
            load[o] += 2.0 ** int(lg)
        assert load.max() / load.mean() < 1.02, (world, load)
    rr = np.zeros(8)   # what the round-robin of round 1 did on eight ranks
    for k, i in enumerate(np.argsort(-logs.astype(np.int64), kind="stable")):
        rr[k % 8] += 2.0 ** int(logs[i])
    assert rr.max() / rr.mean() > 1.3
