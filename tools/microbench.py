#!/usr/bin/env python
"""Integer-pipe micro-benchmark (csrc/microbench.cu) -> JSON: the measured peak the Blake2s roofline is reported against."""
import importlib
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("stwo-brainfuck_b200")
NAMES = ["LOP3", "SHF", "PRMT", "IADD3", "IMAD", "LOP3+IMAD 1:1", "blake2s G mix (8 ALU + 6 IMAD per G)"]


def run(be, iters=8192):
    out = {}
    for k, name in enumerate(NAMES):
        r = be.microbench_int(k, iters)
        r["total_ops_per_clk_sm"] = r["alu_ops_per_clk_sm"] + r["fma_ops_per_clk_sm"]
        out[name] = r
    return out


if __name__ == "__main__":
    be = pkg.CudaBackend(0)
    print(json.dumps(run(be), indent=1))
    be.close()
