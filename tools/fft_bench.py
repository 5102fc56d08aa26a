"""FFT micro-benchmark (one GPU): interpolate and 2x LDE at a few sizes, CUDA-event timed, inputs larger than L2.
usage: [STWO_CUDA_LIB=path/to/variant.so] python tools/fft_bench.py"""
import importlib, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
pkg = importlib.import_module("stwo-brainfuck_b200")
st = torch.cuda.Stream(); torch.cuda.set_stream(st)
be = pkg.CudaBackend(0, st.cuda_stream)
tw = be.precompute_twiddles(26)
out = {"lib": os.path.basename(pkg.LIB_PATH)}
for log, ncols in ((16, 64), (20, 32), (22, 16), (24, 8), (25, 4)):
    host = np.random.default_rng(log).integers(0, pkg.P, size=1 << log, dtype=np.uint32)
    cols = [be.column(host) for _ in range(ncols)]
    e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    for it in range(3):
        e[0].record(st); be.interpolate_columns(cols, tw); e[1].record(st)
        e[2].record(st); ev = be.evaluate_polynomials(cols, 1, tw); e[3].record(st)
        torch.cuda.synchronize()
        for c in ev: c.free()
    gb = ncols * (1 << log) * 4 / 1e9
    ti, tv = e[0].elapsed_time(e[1]), e[2].elapsed_time(e[3])
    out[f"log{log}x{ncols}"] = {"interpolate_ms": round(ti, 4), "interpolate_GBs": round(2 * gb / ti * 1e3, 1),
                               "lde_ms": round(tv, 4), "lde_GBs": round(3 * gb / tv * 1e3, 1)}
    for c in cols: c.free()
print(json.dumps(out))
