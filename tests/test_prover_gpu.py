"""End-to-end: prove on the GPU, verify on the host, and compare the proof byte for byte with the CPU oracle's proof
(same protocol driver, every device op replaced by the scalar restatement).  Mirrors the reference's own end-to-end
tests, crates/brainfuck_prover/src/brainfuck_air/mod.rs:799-859 (prove -> verify on four tiny programs)."""
import ctypes
import os

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PROGRAMS = os.path.join(ROOT, "tests", "golden", "programs")


def oracle_proof(orc, code: bytes, stdin: bytes, log_max_rows: int) -> str:
    lib = orc.lib
    lib.orc_prove_json.restype = ctypes.c_void_p
    lib.orc_last_error.restype = ctypes.c_char_p
    p = lib.orc_prove_json(code, stdin, ctypes.c_size_t(len(stdin)), ctypes.c_uint32(log_max_rows), 1)
    assert p, lib.orc_last_error()
    s = ctypes.string_at(p).decode()
    lib.orc_free(ctypes.c_void_p(p))
    return s


CASES = [  # programs of the kind the reference's end-to-end tests use (mod.rs:804-858; the exact four run on the CPU oracle in
          # tests/test_golden_proofs.py::test_reference_end_to_end_programs) + the shipped examples that finish quickly
    ("with_input", b"+>,<[>+.<-]", b"\x01", 10),
    ("no_input", b"+++>++<[->+<]>.", b"", 10),
    ("jump_mid", b"++[>+<-]>[-]<", b"", 10),
    ("hello_kakarot", None, b"", 17),
    ("collatz", None, b"7\n", 21),
]


@pytest.mark.parametrize("name,code,stdin,lmr", CASES)
def test_prove_verify_matches_oracle(pkg, be, orc, name, code, stdin, lmr):
    if code is None:
        code = open(os.path.join(PROGRAMS, name + ".bf"), "rb").read()
    proof = pkg.prove_brainfuck(be, code, stdin, lmr)
    proof.verify()
    want = oracle_proof(orc, code, stdin, lmr)
    got = proof.json()
    assert len(got) == len(want)
    assert got == want, f"{name}: GPU proof differs from the oracle's proof"


def test_hello_kakarot_at_reference_test_size(pkg, be):
    # LOG_MAX_ROWS = 20 is what the reference uses under cfg(test) (brainfuck_air/mod.rs:430-433)
    code = open(os.path.join(PROGRAMS, "hello_kakarot.bf"), "rb").read()
    proof = pkg.prove_brainfuck(be, code, b"", 20)
    proof.verify()
    assert proof.output() == b"Hello Kakarot World!\n"
    r = proof.report()
    assert r["steps"] == 651 and r["log_sizes"] == [17, 14, 12, 14, 8, 4, 4, 10, 10, 9, 13, 11, 4]


@pytest.mark.parametrize("what", range(8))
def test_verifier_rejects_tampered_proof(pkg, be, what):
    proof = pkg.prove_brainfuck(be, b"+>,<[>+.<-]", b"\x03", 10)
    proof.verify()
    proof.verify_json()                      # the wire text, parsed back and verified with the verifier's own LOG_MAX_ROWS
    bad = proof.tamper(what)                 # corrupted on the wire text, read back through sbf_proof_from_json
    with pytest.raises(pkg.VerificationError):
        bad.verify()
    with pytest.raises(pkg.VerificationError):
        bad.verify_json()
    with pytest.raises(pkg.VerificationError):
        proof.verify_json(11)                # a verifier with another LOG_MAX_ROWS expects another preprocessed tree


def test_component_too_large_is_an_error(pkg, be):
    import gc
    gc.collect()
    code = open(os.path.join(PROGRAMS, "hello_kakarot.bf"), "rb").read()
    before = be.live_columns()
    with pytest.raises(pkg.ProvingError):
        pkg.prove_brainfuck(be, code, b"", 12)   # Memory needs log 17 > LOG_MAX_ROWS 12 (the reference panics likewise)
    assert be.live_columns() == before            # the failed proof released the preprocessed tree it had already built
    pkg.prove_brainfuck(be, code, b"", 17).verify()   # and the context is still usable
    gc.collect()
    assert be.live_columns() == before            # a finished proof leaves nothing behind either


def test_sierpinski_is_outside_the_reference_envelope(pkg, be):
    """BASELINE.json configs[2] lists sierpinski.bf; its Memory table needs log size 29 > LOG_MAX_ROWS 24:
    the reference cannot prove it (no IsFirst(29) column, twiddles too small, brainfuck_air/mod.rs:427-428,480-484) and neither
    may the CUDA path pretend to — a ProvingError naming the component, raised before 1 GiB of table is built, nothing leaked."""
    import gc
    gc.collect()
    code = open(os.path.join(PROGRAMS, "sierpinski.bf"), "rb").read()
    before = be.live_columns()
    for overlap in (True, False):
        with pytest.raises(pkg.ProvingError, match="component too large: memory"):
            pkg.prove_brainfuck(be, code, b"", 24, overlap_host=overlap)
        assert be.live_columns() == before


def test_vm_errors_surface_as_proving_errors(pkg, be):
    for code, stdin in ((b"+]", b""), (b",", b""), (b"<+", b""), (b"", b"")):
        with pytest.raises(pkg.ProvingError):
            pkg.prove_brainfuck(be, code, stdin, 10)
