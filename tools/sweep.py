"""BASELINE.json configs[4]: isolated kernel sweep — circle FFT (interpolate, LDE), Blake2s Merkle commit and FRI folds at
log_size 16..26 against the measured HBM peak (and, for Blake2s, the integer-pipe rate).  Run on the GPU box:
    python tools/sweep.py > gpurun_out/sweep.json
Inputs follow SURVEY.md §8(d) C5: uniformly random u32 < P from numpy default_rng(0x5EED0000 + log_size).
Every timing is CUDA events on the launch stream, best of 5 after 2 warm-ups; inputs at log >= 24 exceed L2 (126 MB)
when batched over the listed column count."""
import importlib
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("stwo-brainfuck_b200")
P = pkg.P
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}
HBM = peaks["hbm_gbs"]

stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
be = pkg.CudaBackend(0, stream.cuda_stream)
tw = be.precompute_twiddles(26)


def timeit(fn, setup=None, reps=5, warm=2):
    best = 1e30
    for i in range(warm + reps):
        arg = setup() if setup else None
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        fn(arg)
        e1.record(stream)
        torch.cuda.synchronize()
        if i >= warm:
            best = min(best, e0.elapsed_time(e1))
    return best


rows = []
for log in range(16, 27):
    ncols = max(1, min(16, (1 << 28) >> log))      # keep the batch at <= 1 GiB of trace words
    rng = np.random.default_rng(0x5EED0000 + log)
    base = [be.column(rng.integers(0, P, size=1 << log, dtype=np.uint32)) for _ in range(ncols)]
    N = 1 << log
    r = {"log_size": log, "columns": ncols}
    # interpolate (in place): 8N bytes per column
    ms = timeit(lambda cols: be.interpolate_columns(cols, tw), setup=lambda: [c.clone() for c in base])
    r["interpolate_ms"] = ms
    r["interpolate_GBs"] = 8 * N * ncols / ms / 1e6
    if log + 1 <= 27:
        coeffs = [c.clone() for c in base]
        be.interpolate_columns(coeffs, tw)
        # LDE to the 2x domain: 12N bytes per column
        ms = timeit(lambda _: be.evaluate_polynomials(coeffs, 1, tw))
        r["lde_ms"] = ms
        r["lde_GBs"] = 12 * N * ncols / ms / 1e6
        del coeffs
    # Merkle commit of ncols equal columns: leaf compressions ceil(ncols/16) per row + 1 per inner node
    ms = timeit(lambda _: be.merkle_commit(base))
    comp = N * ((ncols + 15) // 16) + (N - 1)
    r["merkle_ms"] = ms
    r["merkle_Gcomp_s"] = comp / ms / 1e6
    r["merkle_GBs"] = (4 * N * ncols + 96 * N) / ms / 1e6
    # FRI: fold one secure column (4 coordinate columns) all the way down to 2 values
    coords = base[:4] if ncols >= 4 else [base[0].clone() for _ in range(4)]

    def fold_all(_):
        line = [be.zeros(N // 2) for _ in range(4)]
        be.fold_circle_into_line(line, coords, log, [1, 2, 3, 4], tw)
        lg = log - 1
        while lg > 1:
            line = be.fold_line(line, lg, [1, 2, 3, 4], tw)
            lg -= 1
    ms = timeit(fold_all)
    r["fri_fold_ms"] = ms
    r["fri_fold_GBs"] = (16 * N + 2 * 8 * N + 2 * (24 * (N // 2))) / ms / 1e6   # circle fold + geometric sum of line folds
    for k in ("interpolate_GBs", "lde_GBs", "merkle_GBs", "fri_fold_GBs"):
        if k in r:
            r[k.replace("GBs", "frac_hbm")] = r[k] / HBM
    rows.append(r)
    del base, coords
    print(json.dumps(r), file=sys.stderr)

print(json.dumps({"hbm_peak_GBs": HBM, "peak_source": "MEASURED_PEAKS.json", "rows": rows}, indent=1))
