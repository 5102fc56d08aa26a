"""Blake2s layer-hash micro-benchmark (one GPU).  usage: [STWO_CUDA_LIB=variant.so] python tools/merkle_bench.py"""
import importlib, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
pkg = importlib.import_module("stwo-brainfuck_b200")
st = torch.cuda.Stream(); torch.cuda.set_stream(st)
be = pkg.CudaBackend(0, st.cuda_stream)
out = {"lib": os.path.basename(pkg.LIB_PATH)}
for log, ncols, prev in ((24, 16, False), (24, 4, True), (25, 4, False), (22, 60, True)):
    cols = [be.column(np.random.default_rng(i).integers(0, pkg.P, size=1 << log, dtype=np.uint32)) for i in range(ncols)]
    pv = be.column(np.random.default_rng(99).integers(0, 2**32, size=16 << log, dtype=np.uint32)) if prev else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for it in range(5):
        e0.record(st); h = be.commit_on_layer(log, pv, cols); e1.record(st); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1)); h.free()
    comp = (1 << log) * ((ncols + 15) // 16 + (1 if prev else 0))
    out[f"log{log}_c{ncols}_{'prev' if prev else 'leaf'}"] = {"ms": round(best, 4), "Gcomp_s": round(comp / best / 1e6, 2)}
    for c in cols: c.free()
    if pv: pv.free()
print(json.dumps(out))
