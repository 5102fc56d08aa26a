"""Profiling driver (run under ncu on the GPU box): a few isolated hot kernels at fib19's largest sizes."""
import importlib, sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("stwo-brainfuck_b200")
which = sys.argv[1] if len(sys.argv) > 1 else "all"
be = pkg.CudaBackend(0)
tw = be.precompute_twiddles(26)
rng = np.random.default_rng(0)
LOG = int(os.environ.get("PROF_LOG", "24"))
cols = [be.column(rng.integers(0, pkg.P, size=1 << LOG, dtype=np.uint32)) for _ in range(8)]
for it in range(2):
    if which in ("all", "fft"):
        be.interpolate_columns(cols, tw)
        ldes = be.evaluate_polynomials(cols, 1, tw)
    else:
        ldes = cols
    if which in ("all", "merkle"):
        layers, root = be.merkle_commit(ldes)
be.sync()
print("done")
