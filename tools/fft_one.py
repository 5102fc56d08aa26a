#!/usr/bin/env python
"""One interpolate + one 2x LDE of 4 columns of 2^25 (ncu driver for the FFT kernels)."""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("stwo-brainfuck_b200")
be = pkg.CudaBackend(0)
tw = be.precompute_twiddles(26)
log = int(sys.argv[1]) if len(sys.argv) > 1 else 25
host = np.random.default_rng(log).integers(0, pkg.P, size=1 << log, dtype=np.uint32)
cols = [be.column(host) for _ in range(4)]
for _ in range(2):
    be.interpolate_columns(cols, tw)
    ev = be.evaluate_polynomials(cols, 1, tw)
    be.sync() if hasattr(be, "sync") else None
    for c in ev:
        c.free()
be.close()
