#!/usr/bin/env python
"""Blake2s Merkle commit on the tree shapes of a proof: G compressions / s per shape (CUDA events, best of 5).
SC_MERKLE_GENERIC=1 switches the <= 4-column specialisation off (A/B)."""
import importlib
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("stwo-brainfuck_b200")
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
be = pkg.CudaBackend(0, stream.cuda_stream)
rng = np.random.default_rng(1)
out = {}
for name, log, ncols in (("4 columns, 2^25 rows (FRI / composition)", 25, 4), ("4 columns, 2^22 rows", 22, 4), ("4 columns, 2^18 rows", 18, 4),
                         ("4 columns, 2^14 rows", 14, 4), ("16 columns, 2^24 rows", 24, 16), ("1 column, 2^25 rows", 25, 1)):
    base = be.column(rng.integers(0, pkg.P, size=1 << log, dtype=np.uint32))
    cols = [base] * ncols
    best = 1e30
    for i in range(7):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        layers, root = be.merkle_commit(cols)
        e1.record(stream)
        torch.cuda.synchronize()
        if i >= 2:
            best = min(best, e0.elapsed_time(e1))
        for l in layers:
            l.free()
    comp = (1 << log) * ((ncols + 15) // 16) + (1 << log) - 1
    out[name] = {"ms": best, "Gcomp_s": comp / best / 1e6}
    base.free()
print(json.dumps(out, indent=1))
be.close()
