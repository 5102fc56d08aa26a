"""The preprocessed tree (IsFirst columns of log size LOG_MAX_ROWS..4) is the same for every program
(crates/brainfuck_prover/src/brainfuck_air/mod.rs:453-464,493-500); a prover may keep it between proofs.
The proofs must not change.  CPU: the driver's cache logic on the oracle backend.  GPU: SBF_CACHE_PREPROCESSED."""
import ctypes
import hashlib
import json
import os

import pytest

import proof_canon

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "proof_hashes.json")))


def source(name):
    g = GOLD[name]
    return g["code"].encode() if g["code"] else open(os.path.join(ROOT, "tests", "golden", "programs", name + ".bf"), "rb").read()


def check(name, js: bytes):
    g = GOLD[name]
    proof_canon.check(js, g)


def test_oracle_driver_with_cache_reproduces_golden_proofs(orc):
    """Same LOG_MAX_ROWS twice (a hit), then another size (rebuild), then the first size again (rebuild)."""
    names = ["with_input", "no_input", "a-bc", "jump_mid"]
    n = len(names)
    codes = (ctypes.c_char_p * n)(*[source(x) for x in names])
    stdin = [bytes.fromhex(GOLD[x]["stdin_hex"]) for x in names]
    bufs = [ctypes.create_string_buffer(s, max(1, len(s))) for s in stdin]
    inputs = (ctypes.c_void_p * n)(*[ctypes.cast(b, ctypes.c_void_p) for b in bufs])
    lens = (ctypes.c_size_t * n)(*[len(s) for s in stdin])
    logs = (ctypes.c_uint32 * n)(*[GOLD[x]["log_max_rows"] for x in names])
    lib = orc.lib
    lib.orc_prove_sequence_cached_json.restype = ctypes.c_void_p
    p = lib.orc_prove_sequence_cached_json(n, codes, inputs, lens, logs, 1)
    assert p, ctypes.string_at(lib.orc_last_error())
    lines = ctypes.string_at(p).split(b"\n")
    lib.orc_free(ctypes.c_void_p(p))
    assert len(lines) == n + 1
    for name, js in zip(names, lines):
        check(name, js)
    assert [GOLD[x]["log_max_rows"] for x in names] == [10, 10, 12, 10]
    assert lines[-1] == b"3 1"   # fills, hits


@pytest.mark.gpu
def test_cuda_prover_with_preprocessed_cache(pkg, be):
    lib = be._lib
    lib.sc_ctx_live_columns.restype = ctypes.c_uint64
    live = lambda: int(lib.sc_ctx_live_columns(be._ctx))
    base = live()
    seq = ["with_input", "no_input", "a-bc", "jump_mid", "hello_kakarot", "hello_kakarot"]
    held = None
    for i, name in enumerate(seq):
        g = GOLD[name]
        proof = pkg.prove_brainfuck(be, source(name), bytes.fromhex(g["stdin_hex"]), g["log_max_rows"], cache_preprocessed=True)
        proof.verify()
        check(name, proof.json().encode())
        now = live() - base
        assert now > 0                     # the tree stays on the context ...
        if i and GOLD[seq[i - 1]]["log_max_rows"] == g["log_max_rows"]:
            assert now == held             # ... and a hit adds nothing to it
        held = now
    # a proof for another LOG_MAX_ROWS drops the cached tree, builds its own, then fails (component larger than LOG_MAX_ROWS)
    g = GOLD["hello_kakarot"]
    with pytest.raises(pkg.ProvingError):
        pkg.prove_brainfuck(be, source("hello_kakarot"), b"", 10, cache_preprocessed=True)
    # the new tree went with the failed proof before it was cached: nothing may be left behind
    assert live() == base
    proof = pkg.prove_brainfuck(be, source("hello_kakarot"), b"", g["log_max_rows"], cache_preprocessed=True)
    check("hello_kakarot", proof.json().encode())
    assert live() - base == held
    with pytest.raises(pkg.ProvingError):  # fails on a hit (fib19 needs LOG_MAX_ROWS 24): the tree was made by an earlier call and stays
        pkg.prove_brainfuck(be, source("fib19"), b"", g["log_max_rows"], cache_preprocessed=True)
    assert live() - base == held
    check("hello_kakarot", pkg.prove_brainfuck(be, source("hello_kakarot"), b"", g["log_max_rows"], cache_preprocessed=True).json().encode())
    # without the flag the cache is neither used nor touched
    check("no_input", pkg.prove_brainfuck(be, source("no_input"), b"", 10).json().encode())
    assert live() - base == held
    pkg.clear_preprocessed_cache(be)
    assert live() == base
