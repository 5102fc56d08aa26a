#!/usr/bin/env python
"""bench.py — measures the proving hot path on B200.

Workloads (config.workload):
  fib19_commit   the "LDE + commit" half of BASELINE.json's metric on fib19.bf's main-trace tree shape (128 columns,
                 log sizes from SURVEY.md Table S): interpolate -> evaluate(blowup 2x) -> Blake2s Merkle commit.
                 A step = one pass over one synthetic batch of that shape.  value = algorithmic GB/s.
Contract: one JSON line on stdout from rank 0 (see task statement): metric/value/unit, e2e through the C ABI with host
buffers, roofline of the dominant kernel group, cpu_baseline (oracle port timed on this box's cores), clocks.
`--impl reference` times the CPU oracle port (the reference itself is Rust + an un-vendored git dependency and cannot be
built in this image) on the same config.
"""
import argparse
import ctypes
import importlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
P = (1 << 31) - 1

# fib19.bf main-trace tree: (log_size, n_columns) per component, in commit order (col. fib19)
FIB19_MAIN = [(24, 8), (22, 8), (11, 4), (22, 9), (19, 13), (11, 13), (4, 11), (20, 11), (19, 11), (4, 11), (20, 11),
              (20, 11), (4, 7)]
ROOT_LOG = 26  # brainfuck_air/mod.rs:480-484: twiddles for CanonicCoset(24+1+2).circle_domain().half_coset


def tree_shape(scale_down=0):
    return [(max(4, lg - scale_down), n) for lg, n in FIB19_MAIN]


def commit_bytes(shape):
    """Algorithmic bytes of one LDE+commit pass: per column of N words: iFFT 8N + LDE 12N;
    Merkle: 4 B per LDE cell read + 32 B per node written + 64 B children read per non-leaf-layer node."""
    fft = sum(n * (8 + 12) * (1 << lg) for lg, n in shape)
    max_lde = max(lg for lg, _ in shape) + 1
    merkle = sum(n * 4 * (2 << lg) for lg, n in shape)
    merkle += sum((32 + (64 if k < max_lde else 0)) * (1 << k) for k in range(max_lde + 1))
    return fft, merkle


def n_compressions(shape):
    max_lde = max(lg for lg, _ in shape) + 1
    per_layer = {}
    for lg, n in shape:
        per_layer[lg + 1] = per_layer.get(lg + 1, 0) + n
    tot = 0
    for k in range(max_lde + 1):
        tot += (1 << k) * ((per_layer.get(k, 0) + 15) // 16 + (1 if k < max_lde else 0))
    return tot


class ClockSampler:
    def __init__(self, dev=0):
        self.rows, self.stop = [], False
        self.dev = dev
        self.t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self.stop:
            try:
                o = subprocess.run(["nvidia-smi", "-i", str(self.dev), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                   capture_output=True, text=True, timeout=5).stdout.strip()
                if o:
                    self.rows.append([x.strip() for x in o.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def __enter__(self):
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop = True
        self.t.join(timeout=6)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(int(float(r[0])) for r in self.rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[3 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": int(float(self.rows[0][1])), "reasons": reasons,
                "samples": len(sm)}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)).get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def oracle_lib():
    lib = ctypes.CDLL(os.path.join(ROOT, "oracle", "liborc.so"))
    return lib


def cpu_commit_sample(scale_down, root_log):
    """Oracle port of the same LDE+commit pass on a tree scaled down by 2^scale_down rows; returns (GB/s, seconds, threads)."""
    lib = oracle_lib()
    u32p = ctypes.POINTER(ctypes.c_uint32)
    shape = tree_shape(scale_down)
    rng = np.random.default_rng(1)
    cols = [rng.integers(0, P, size=1 << lg, dtype=np.uint32) for lg, n in shape for _ in range(n)]
    lib.orc_precompute_twiddles(root_log, np.empty(1 << root_log, dtype=np.uint32).ctypes.data_as(u32p),
                                np.empty(1 << root_log, dtype=np.uint32).ctypes.data_as(u32p))  # warm the tree cache
    t0 = time.perf_counter()
    ldes = []
    for c in cols:
        lg = int(np.log2(c.size))
        lib.orc_interpolate(c.ctypes.data_as(u32p), lg, 1, root_log)
        o = np.empty(2 << lg, dtype=np.uint32)
        lib.orc_evaluate(c.ctypes.data_as(u32p), lg, 1, 1, root_log, o.ctypes.data_as(u32p))
        ldes.append(o)
    logs = np.array([int(np.log2(o.size)) for o in ldes], dtype=np.uint32)
    total = sum(8 << k for k in range(int(logs.max()) + 1))
    buf = np.empty(total, dtype=np.uint32)
    ptrs = (u32p * len(ldes))(*[o.ctypes.data_as(u32p) for o in ldes])
    lib.orc_merkle_commit(ptrs, logs.ctypes.data_as(u32p), len(ldes), buf.ctypes.data_as(u32p))
    dt = time.perf_counter() - t0
    fft, merkle = commit_bytes(shape)
    return (fft + merkle) / dt / 1e9, dt, int(lib.orc_num_threads())


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sd = 6  # bounded sample: the same tree with 2^6 fewer rows per column (~15 M cells)
    root_log = ROOT_LOG - sd
    vals, secs = [], []
    for i in range(args.warmup + args.steps):
        v, dt, thr = cpu_commit_sample(sd, root_log)
        if i >= args.warmup:
            vals.append(v)
            secs.append(dt)
    v = float(np.mean(vals))
    line = {"impl": "reference", "metric": "LDE+commit throughput", "value": v, "unit": "GB/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean(secs)), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u32 (M31)", "data": "synthetic",
            "config": {"workload": "fib19_commit", "sample": f"fib19 main-trace tree shape with 2^{sd} fewer rows per column"},
            "cpu_baseline": {"value": v, "unit": "GB/s", "cores": thr, "kind": "port",
                             "sample": f"oracle port (reference is Rust, not buildable here), tree / 2^{sd}"},
            "e2e": {"value": v, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda")
    ap.add_argument("--scale-down", type=int, default=0, help="debug: shrink every column by 2^k rows")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    pkg = importlib.import_module("stwo-brainfuck_b200")
    stream = torch.cuda.Stream()          # a real (non-default) stream: the library launches on this handle,
    torch.cuda.set_stream(stream)         # and torch.cuda.Event records on the same one
    assert stream.cuda_stream != 0
    be = pkg.CudaBackend(local, stream.cuda_stream)
    shape = tree_shape(args.scale_down)
    root_log = ROOT_LOG - args.scale_down
    tw = be.precompute_twiddles(root_log)

    # synthetic trace columns: pinned host buffers (e2e) and resident device copies (value)
    rng = np.random.default_rng(0x5EED0000 + rank)
    host = []
    for lg, n in shape:
        for _ in range(n):
            t = torch.from_numpy(rng.integers(0, P, size=1 << lg, dtype=np.int64).astype(np.int32)).pin_memory()
            host.append(t)
    h2d = sum(t.numel() * 4 for t in host)
    resident = [be.column(t.numpy().view(np.uint32)) for t in host]

    def step_resident():
        cols = [c.clone() for c in resident]       # interpolate is in place; the clone is outside the algorithmic bytes
        be.interpolate_columns(cols, tw)
        ldes = be.evaluate_polynomials(cols, 1, tw)
        layers, root = be.merkle_commit(ldes)
        return root

    def step_e2e():
        cols = [be.column(t.numpy().view(np.uint32)) for t in host]
        be.interpolate_columns(cols, tw)
        ldes = be.evaluate_polynomials(cols, 1, tw)
        layers, root = be.merkle_commit(ldes)
        return root

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = be.launch_count()
        e0.record(stream)
        for _ in range(steps):
            root = fn()
        e1.record(stream)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms / steps, be.launch_count() - l0, root

    fft_b, merkle_b = commit_bytes(shape)
    alg = fft_b + merkle_b
    with ClockSampler(local) as cs:
        ms, launches, root = timed(step_resident, args.steps, args.warmup)
        ms_e2e, _, root2 = timed(step_e2e, max(1, args.steps // 2), 1)
    assert (root == root2).all()
    clocks = cs.summary()

    # per-stage device times (CUDA events on the launch stream) for the roofline of the dominant stage
    def stage_times():
        cols = [c.clone() for c in resident]
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        torch.cuda.synchronize()
        ev[0].record(stream)
        be.interpolate_columns(cols, tw)
        ev[1].record(stream)
        ldes = be.evaluate_polynomials(cols, 1, tw)
        ev[2].record(stream)
        be.merkle_commit(ldes)
        ev[3].record(stream)
        torch.cuda.synchronize()
        return [ev[i].elapsed_time(ev[i + 1]) for i in range(3)]

    st = np.array([stage_times() for _ in range(3)]).min(axis=0)
    peak, peak_src = peaks()
    cells = sum(n << lg for lg, n in shape)
    fft_ms = float(st[0] + st[1])
    roof = {"bound": "hbm", "kernel": "fft_pass_kernel (interpolate + evaluate, all passes)",
            "achieved": fft_b / (fft_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s", "traffic": None, "peak_source": peak_src}
    roof["frac"] = roof["achieved"] / peak

    line = {"metric": "LDE+commit throughput", "value": world * alg / (ms * 1e-3) / 1e9, "unit": "GB/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u32 (M31)", "data": "synthetic",
            "config": {"workload": "fib19_commit", "columns": sum(n for _, n in shape), "trace_cells": cells,
                       "max_log_size": max(lg for lg, _ in shape), "log_blowup": 1, "scale_down": args.scale_down,
                       "l2": "inputs (1 GB per step) exceed L2", "parallelism": f"replicas x{world}"},
            "e2e": {"value": world * alg / (ms_e2e * 1e-3) / 1e9, "unit": "GB/s", "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 32},
            "gpu_launches": int(launches),
            "stages_ms": {"interpolate": float(st[0]), "evaluate": float(st[1]), "merkle_commit": float(st[2])},
            "merkle": {"compressions": n_compressions(shape), "gcomp_per_s": n_compressions(shape) / (float(st[2]) * 1e-3) / 1e9},
            "roofline": roof, "clocks": clocks}
    if rank == 0 and not args.no_cpu_baseline:
        sd = 6
        v, dt, thr = cpu_commit_sample(sd, ROOT_LOG - sd)
        line["cpu_baseline"] = {"value": v, "unit": "GB/s", "cores": thr, "kind": "port", "seconds": dt,
                                "sample": f"oracle port of the same pass on the tree with 2^{sd} fewer rows per column"}
    if rank == 0:
        print(json.dumps(line))
    be.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
