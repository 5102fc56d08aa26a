#!/usr/bin/env python
"""Writes the CUDA prover's wire-format proofs of the BASELINE.json programs at LOG_MAX_ROWS 24 to gpurun_out/cuda_proofs/ —
the files tools/make_reference_goldens.sh feeds to the reference's own `brainfuck_prover verify`."""
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("stwo-brainfuck_b200")
out = os.path.join(ROOT, "gpurun_out", "cuda_proofs")
os.makedirs(out, exist_ok=True)
be = pkg.CudaBackend(0)
for name, stdin in (("hello_kakarot", b""), ("fib19", b""), ("collatz", b"7\n"), ("synthetic_2p24", b"")):
    code = open(os.path.join(ROOT, "tests", "golden", "programs", name + ".bf"), "rb").read()
    pr = pkg.prove_brainfuck(be, code, stdin, 24)
    pr.verify()
    open(os.path.join(out, name + ".proof.json"), "w").write(pr.json())
    print(name, len(pr.json()), "bytes")
be.close()
