#!/usr/bin/env python
"""Proves a program N times on one GPU and prints the kernel-launch count of every proof (ncu / compute-sanitizer driver).
usage: python tools/prove_once.py [program=fib19] [n=2] [log_max_rows=24]"""
import importlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("stwo-brainfuck_b200")
name = sys.argv[1] if len(sys.argv) > 1 else "fib19"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 2
lmr = int(sys.argv[3]) if len(sys.argv) > 3 else 24
stdin = b"7\n" if name == "collatz" else b""
code = open(os.path.join(ROOT, "tests", "golden", "programs", name + ".bf"), "rb").read()
be = pkg.CudaBackend(0)
counts = []
for _ in range(n):
    l0 = be.launch_count()
    pr = pkg.prove_brainfuck(be, code, stdin, lmr, overlap_host=False)
    counts.append(be.launch_count() - l0)
pr.verify()
print(json.dumps({"program": name, "log_max_rows": lmr, "launches_per_proof": counts}))
be.close()
