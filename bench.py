#!/usr/bin/env python
"""bench.py — measures the proving hot path on B200.

Workloads (config.workload):
  fib19_prove   (default) BASELINE.json configs[1]: prove fib19.bf end to end on one B200, LOG_MAX_ROWS = 24, PcsConfig
                default (pow 5, blowup 2x, 3 queries).  A step = one full proof.  metric = prove time (s), lower is better.
                value = device path (tables already built on the host; includes the 63 MB compact-column upload);
                e2e   = the whole `prove` call a user makes: VM run + host table building + uploads + proof + proof readback.
  fib19_commit  the "LDE + commit GB/s" half of the metric on fib19's main-trace tree shape (128 columns): interpolate ->
                evaluate(2x) -> Blake2s Merkle commit.  value = algorithmic GB/s.  (--workload commit)
Contract: one JSON line on stdout from rank 0: metric/value/unit, e2e, roofline of the dominant kernel class (device time
from CUDA events on the launch stream, recorded by the library's profiling scopes), cpu_baseline, clocks, gpu_launches.
`--impl reference`: the reference is Rust + an un-vendored git dependency and cannot be built in this image, so this arm
times the in-repo CPU oracle prover (OpenMP, all host cores) on a bounded sample and scales it to the workload (see
`cpu_baseline.sample`).
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
PROGRAMS = os.path.join(ROOT, "tests", "golden", "programs")

# fib19.bf: (log_size, main columns, LogUp columns) per component in commit order (column fib19)
FIB19 = [(24, 8, 1), (22, 8, 1), (11, 4, 1), (22, 9, 3), (19, 13, 1), (11, 13, 1), (4, 11, 1), (20, 11, 1), (19, 11, 1),
         (4, 11, 1), (20, 11, 1), (20, 11, 1), (4, 7, 1)]
COLLATZ = [(21, 8, 1), (17, 8, 1), (13, 4, 1), (17, 9, 3), (14, 13, 1), (13, 13, 1), (5, 11, 1), (15, 11, 1), (14, 11, 1),
           (6, 11, 1), (14, 11, 1), (15, 11, 1), (4, 7, 1)]
ROOT_LOG = 26  # brainfuck_air/mod.rs:480-484: twiddles for CanonicCoset(24+1+2).circle_domain().half_coset


def proof_columns(shape, log_max_rows):
    """log sizes of every committed polynomial of a proof: preprocessed, main, interaction, composition."""
    pre = list(range(log_max_rows, 3, -1))
    main = [lg for lg, n, _ in shape for _ in range(n)]
    inter = [lg for lg, _, k in shape for _ in range(4 * k)]
    comp = [max(lg for lg, _, _ in shape) + 1] * 4
    return pre, main, inter, comp


def fft_bytes(logs):
    """algorithmic bytes of interpolate (8N) + evaluate on the 2x domain (12N) per polynomial of 2^log words"""
    return sum((8 + 12) * (1 << lg) for lg in logs)


def proof_fft_bytes(shape, log_max_rows):
    pre, main, inter, comp = proof_columns(shape, log_max_rows)
    # composition polynomials are interpolated once (accumulator finalize) and evaluated once on the 2x domain; the lifts of
    # the running polynomial between accumulator sizes are small and not counted
    return fft_bytes(pre) + fft_bytes(main) + fft_bytes(inter) + fft_bytes(comp)


def tree_stats(lde_logs, rep=0):
    """(algorithmic bytes, Blake2s compressions) of MerkleProver::commit over columns of the given LDE log sizes: a layer of R
    rows with C injected columns moves R*(4C + 32 + 64*[has children]) bytes and takes R*(ceil(C/16) + [has children])
    compressions.  rep > 0: what the kernels EXECUTE when every column repeats each value 2^rep times (the
    deepest rep layers hash one node per group)."""
    from collections import Counter
    by, top = Counter(lde_logs), max(lde_logs)
    nbytes = comps = 0
    for k in range(top, -1, -1):
        c, prev = by.get(k, 0), k < top
        rows = 1 << k
        nbytes += rows * (4 * c + 32 + (64 if prev else 0))
        comps += (rows >> max(0, rep - (top - k))) * (-(-c // 16) + (1 if prev else 0))
    return nbytes, comps


def proof_merkle_stats(shape, log_max_rows):
    """All trees of one proof: the four commitment trees, the FRI first layer (4 coordinate columns per distinct LDE size)
    and the FRI inner layers (line evaluations of 2^(top-1) ... 2^2, 4 columns each; last-layer degree bound 0, blow-up 2x).
    Returns (algorithmic bytes, algorithmic compressions, executed compressions)."""
    pre, main, inter, comp = proof_columns(shape, log_max_rows)
    trees = [([l + 1 for l in t], 0) for t in (pre, inter, comp)] + [([l + 1 for l in main], 4)]
    sizes = sorted({l + 1 for l in pre + main + inter + comp})
    trees.append(([l for l in sizes for _ in range(4)], 0))
    trees += [([l] * 4, 0) for l in range(max(sizes) - 1, 1, -1)]
    nbytes = comps = done = 0
    for logs, rep in trees:
        b, c = tree_stats(logs)
        nbytes += b
        comps += c
        done += tree_stats(logs, rep)[1]
    return nbytes, comps, done


ALU_OPS_PER_COMPRESSION = 8 * 80 + 24  # per G: 4 XOR + 4 rotates stay on the ALU pipe (the 6 additions run as IMAD); + finalisation


def proof_lde_cells(shape, log_max_rows):
    pre, main, inter, comp = proof_columns(shape, log_max_rows)
    return sum(2 << lg for lg in pre + main + inter + comp)


def commit_bytes(shape):
    cols = [(lg, n) for lg, n, _ in shape]
    fft = sum(n * (8 + 12) * (1 << lg) for lg, n in cols)
    max_lde = max(lg for lg, _ in cols) + 1
    merkle = sum(n * 4 * (2 << lg) for lg, n in cols)
    merkle += sum((32 + (64 if k < max_lde else 0)) * (1 << k) for k in range(max_lde + 1))
    return fft, merkle


class ClockSampler:
    def __init__(self, dev=0):
        self.rows, self.stop, self.dev = [], False, dev
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
            time.sleep(0.1)

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
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": int(float(self.rows[0][1])), "reasons": reasons, "samples": len(sm)}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)).get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def rooflines(kern, shape, lmr, world, clocks):
    """`roofline`: the dominant kernel class of a proof, commit_layer_kernel (Blake2s Merkle layers, about half of the kernel
    time), per rank: algorithmic bytes / summed CUDA-event time against the measured HBM peak, as the contract asks, plus the
    bound that actually binds it (`alu_pipe`: compressions/s against SMs x 4 SMSPs x 16 ALU lanes x SM clock / ALU ops per
    compression).  `roofline_fft`: the same for the FFT launches (second largest class)."""
    peak, peak_src = peaks()
    mb, mc, mx = proof_merkle_stats(shape, lmr)
    m_ms = kern.get("merkle_commit_layer", 0)
    share = m_ms / sum(kern.values()) if kern else None
    sm_hz = (clocks.get("sm_mhz") or 1965) * 1e6
    bound = 148 * 4 * 16 * sm_hz / ALU_OPS_PER_COMPRESSION
    roof = {"bound": "hbm", "kernel": "commit_layer_kernel + commit_top_kernel (every Merkle launch of one proof, this rank's share)",
            "achieved": mb / world / (m_ms * 1e-3) / 1e9 if m_ms else None, "peak": peak, "unit": "GB/s", "traffic": None,
            "peak_source": peak_src, "algorithmic_bytes": mb / world, "ms_per_proof": m_ms, "share_of_kernel_time": share,
            "traffic_note": "ncu --set full on seven launches of a fib19 proof: dram read+write = 0.96-1.00 x the algorithmic bytes "
                            "of the launch (profiles/r1_final_ncu_full_merkle_*.csv); not summed over the ~170 launches, hence null",
            "alu_pipe": {"algorithmic_compressions": mc / world, "executed_compressions": mx / world,
                         "achieved_Gcomp_s": mx / world / (m_ms * 1e-3) / 1e9 if m_ms else None, "bound_Gcomp_s": bound / 1e9,
                         "frac": (mx / world / (m_ms * 1e-3)) / bound if m_ms else None,
                         "alu_ops_per_compression": ALU_OPS_PER_COMPRESSION, "ncu_alu_pipe_active_pct": "77-82 (profiles/r1_final_ncu_full_merkle_*.csv)",
                         "note": "the binding roofline: XOR/rotate run on the 16-lane ALU pipe; executed < algorithmic because "
                                 "the main-trace tree hashes one node per 16 repeated rows in its four deepest layers"}}
    roof["frac"] = roof["achieved"] / peak if roof["achieved"] else None
    fft_ms = kern.get("fft_interpolate", 0) + kern.get("fft_evaluate", 0)
    fb = proof_fft_bytes(shape, lmr) / world
    roof_fft = {"bound": "hbm", "kernel": "fft_kernel (every interpolate + evaluate launch of one proof, this rank's share)",
                "achieved": fb / (fft_ms * 1e-3) / 1e9 if fft_ms else None, "peak": peak, "unit": "GB/s", "traffic": None,
                "algorithmic_bytes": fb, "ms_per_proof": fft_ms,
                "note": "algorithmic bytes count every column at full length (8N + 12N); main-trace columns are transformed on "
                        "their 2^-4 distinct values",
                "binding_limit": "integer issue: ncu smsp__issue_active 59-72 %, ALU pipe 47-61 %, DRAM 11-39 % on the passes of a "
                                 "fib19 proof (profiles/r1_final_ncu_full_fft_a.csv); >= 10 integer instructions per butterfly"}
    roof_fft["frac"] = roof_fft["achieved"] / peak if roof_fft["achieved"] else None
    return roof, roof_fft


# ------------------------------------------------------------------------------------------------ CPU baseline (oracle port)
def oracle():
    lib = ctypes.CDLL(os.path.join(ROOT, "oracle", "liborc.so"))
    lib.orc_prove_json.restype = ctypes.c_void_p
    lib.orc_last_error.restype = ctypes.c_char_p
    return lib


def cpu_prove_sample():
    """CPU oracle prover (all host threads) on collatz.bf (stdin '7\\n') at LOG_MAX_ROWS 21, scaled to fib19 by committed LDE
    cells.  Returns (scaled seconds, measured seconds, threads, description)."""
    lib = oracle()
    code = open(os.path.join(PROGRAMS, "collatz.bf"), "rb").read()
    t0 = time.perf_counter()
    p = lib.orc_prove_json(code, b"7\n", ctypes.c_size_t(2), ctypes.c_uint32(21), 0)
    dt = time.perf_counter() - t0
    if not p:
        raise RuntimeError(lib.orc_last_error())
    lib.orc_free(ctypes.c_void_p(p))
    scale = proof_lde_cells(FIB19, 24) / proof_lde_cells(COLLATZ, 21)
    return dt * scale, dt, int(lib.orc_num_threads()), \
        f"CPU oracle prover (port; the Rust reference is not buildable here) on collatz.bf, LOG_MAX_ROWS 21: {dt:.2f} s measured, " \
        f"scaled x{scale:.2f} by committed LDE cells to fib19.bf at LOG_MAX_ROWS 24"


def run_reference(args):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    vals = []
    for i in range(max(1, min(args.steps, 2))):  # each sample is ~15-30 s of CPU work
        scaled, dt, thr, desc = cpu_prove_sample()
        vals.append(scaled)
    v = float(np.mean(vals))
    line = {"impl": "reference", "metric": "fib19.bf prove time", "value": v, "unit": "s", "n_gpus": args.gpus,
            "steps": len(vals), "warmup": 0, "ms_per_step": 1e3 * v, "higher_is_better": False, "scaling": "weak",
            "vs_baseline": None, "dtype": "u32 (M31)", "data": "fib19.bf (reference example program)",
            "config": {"workload": "fib19_prove", "log_max_rows": 24, "pcs": "pow 5, blowup 2x, 3 queries"},
            "cpu_baseline": {"value": v, "unit": "s", "cores": thr, "kind": "port", "sample": desc},
            "e2e": {"value": v, "unit": "s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------ GPU arms
def setup(args):
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
    stream = torch.cuda.Stream()   # a real (non-default) stream: the library launches on this handle and
    torch.cuda.set_stream(stream)  # torch.cuda.Event records on the same one
    be = pkg.CudaBackend(local, stream.cuda_stream)
    return torch, dist, rank, world, local, pkg, stream, be


def max_over_ranks(torch, dist, world, x):
    if world == 1:
        return x
    t = torch.tensor([x], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def bench_prove_sharded(args, torch, dist, rank, world, local, pkg, stream, be):
    """N > 1: ONE fib19 proof split over the N GPUs (strong scaling) by the sharded prover: column-sharded FFTs, one NCCL
    all-to-all per commitment tree, row-sharded hashing / constraints / quotients / FRI (csrc/host/prover_sharded.hpp)."""
    import hashlib
    code = open(os.path.join(PROGRAMS, "fib19.bf"), "rb").read()
    lmr = 24
    comm = pkg.Comm.from_torch_distributed(be, dist)
    for _ in range(args.warmup):
        pkg.prove_brainfuck_sharded(be, comm, code, b"", lmr)
    dist.barrier()
    torch.cuda.synchronize()
    be.profile(True)
    be.profile_report()
    l0 = be.launch_count()
    reports = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as cs:
        t0 = time.perf_counter()
        e0.record(stream)
        for _ in range(args.steps):
            pr = pkg.prove_brainfuck_sharded(be, comm, code, b"", lmr)
            reports.append(pr.report())
            js = pr.json()                  # proof readback inside the timed region
        e1.record(stream)
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
    dist.barrier()
    launches = be.launch_count() - l0
    prof = be.profile_report()
    be.profile(False)
    pr.verify()
    digest = hashlib.sha256(js.encode()).hexdigest()
    digests = [None] * world
    dist.all_gather_object(digests, digest)
    assert len(set(digests)) == 1, "ranks disagree on the proof"
    e2e_s = max_over_ranks(torch, dist, world, max(wall / args.steps, e0.elapsed_time(e1) * 1e-3 / args.steps))
    # each table is built by one rank; the others wait for the slowest builder at the first collective, so the host share of
    # a proof is the MAX over ranks of the table-building time (rank 0 builds the Memory table, the longest one)
    host_tables = max_over_ranks(torch, dist, world, float(np.mean([r["stages_ms"]["tables(host)"] for r in reports])) * 1e-3)
    dev_s = max_over_ranks(torch, dist, world, float(np.mean([r["prove_ms"] for r in reports])) * 1e-3) - host_tables
    stages = {k: float(np.mean([r["stages_ms"][k] for r in reports])) for k in reports[0]["stages_ms"]}
    kern = {k: v[0] / args.steps for k, v in sorted(prof.items(), key=lambda kv: -kv[1][0])}
    clocks = cs.summary()
    roof, roof_fft = rooflines(kern, FIB19, lmr, world, clocks)
    h2d = sum(n * (1 << (lg - 4)) * 4 for lg, n, _ in FIB19)
    line = {"metric": "fib19.bf prove time", "value": dev_s, "unit": "s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_s * 1e3, "higher_is_better": False, "scaling": "strong", "vs_baseline": None, "dtype": "u32 (M31)",
            "data": "fib19.bf (reference example program), 199246 VM steps",
            "config": {"workload": "fib19_prove", "log_max_rows": lmr, "pcs": "pow 5, blowup 2x, 3 queries", "columns": 213,
                       "lde_cells": proof_lde_cells(FIB19, lmr),
                       "parallelism": f"one proof over {world} GPUs: column-sharded FFT -> all_to_all per tree -> row-sharded "
                                      "Merkle / constraints / quotients / FRI; sub-roots all-gathered; main-trace tree from "
                                      "local transforms of the replicated compact columns (no exchange)",
                       "proof_sha256": digest},
            "e2e": {"value": e2e_s, "unit": "s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": len(js),
                    "includes": "VM run and host table building on every rank, uploads, proof, proof readback"},
            "gpu_launches": int(launches // args.steps), "stages_ms": stages, "kernel_ms_per_proof": kern, "roofline": roof,
            "roofline_fft": roof_fft, "clocks": clocks, "verified": True}
    if args.timeline:  # one more proof with the scopes kept in launch order: where this rank's device sits idle
        dist.barrier()
        be.profile(True); be.profile_report()
        pr = pkg.prove_brainfuck_sharded(be, comm, code, b"", lmr)
        tl = be.profile_timeline()
        be.profile_report(); be.profile(False)
        json.dump({"rank": rank, "world": world, "timeline": tl, "stages_ms": pr.report()["stages_ms"]},
                  open(f"{args.timeline}.rank{rank}.json", "w"))
    if rank == 0:
        print(json.dumps(line))
    comm.close()
    be.close()
    dist.destroy_process_group()


def bench_prove(args):
    torch, dist, rank, world, local, pkg, stream, be = setup(args)
    if world > 1:
        return bench_prove_sharded(args, torch, dist, rank, world, local, pkg, stream, be)
    code = open(os.path.join(PROGRAMS, "fib19.bf"), "rb").read()
    lmr = 24

    for _ in range(args.warmup):
        pkg.prove_brainfuck(be, code, b"", lmr)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    # (A) device path: host tables built first (no overlap), value = prove time minus the host table building
    be.profile(True)
    be.profile_report()
    l0 = be.launch_count()
    reports = []
    with ClockSampler(local) as cs:
        for _ in range(args.steps):
            reports.append(pkg.prove_brainfuck(be, code, b"", lmr, overlap_host=False).report())
        launches = be.launch_count() - l0
        prof = be.profile_report()
        be.profile(False)
        # (B) end to end, the call a user makes: VM + tables (overlapped with phase 0) + uploads + proof + proof readback
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(stream)
        for _ in range(args.steps):
            pr = pkg.prove_brainfuck(be, code, b"", lmr)
            proof_len = len(pr.json())      # proof readback (D2H of the result) is inside the timed region
        e1.record(stream)
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        # (C) not the headline: the same end-to-end call with the program-independent preprocessed tree kept on the context
        # between proofs (SBF_CACHE_PREPROCESSED).  The reference rebuilds that tree in every proof,
        # so `value` and `e2e` above do too; this leg only reports what a server proving many programs would see.
        cached = None
        try:
            pkg.prove_brainfuck(be, code, b"", lmr, cache_preprocessed=True)   # fills the cache
            torch.cuda.synchronize()
            t0c = time.perf_counter()
            for _ in range(args.steps):
                prc = pkg.prove_brainfuck(be, code, b"", lmr, cache_preprocessed=True)
                same = prc.json() == pr.json()
            torch.cuda.synchronize()
            cached = {"value": (time.perf_counter() - t0c) / args.steps, "unit": "s", "proof_identical": bool(same),
                      "note": "end to end as e2e, preprocessed tree reused across proofs; not the headline"}
            pkg.clear_preprocessed_cache(be)
        except Exception as e:  # reported, never hidden: the headline legs above do not depend on this one
            cached = {"error": repr(e)}
        # (D) the strict reading of the reference's prove_brainfuck: nothing kept between proofs, the twiddle tree of
        # half_odds(26) recomputed in every proof as well (brainfuck_air/mod.rs:480-484; SBF_NO_TWIDDLE_CACHE).
        strict = None
        try:
            pkg.prove_brainfuck(be, code, b"", lmr, twiddle_cache=False)
            torch.cuda.synchronize()
            t0s = time.perf_counter()
            for _ in range(args.steps):
                prs = pkg.prove_brainfuck(be, code, b"", lmr, twiddle_cache=False)
                same_s = prs.json() == pr.json()
            torch.cuda.synchronize()
            strict = {"value": (time.perf_counter() - t0s) / args.steps, "unit": "s", "proof_identical": bool(same_s),
                      "note": "end to end as e2e with the twiddle tree recomputed in every proof, as the reference does"}
        except Exception as e:
            strict = {"error": repr(e)}
    if world > 1:
        dist.barrier()
    pr.verify()                              # the host verifier accepts the last proof
    ev_s = e0.elapsed_time(e1) * 1e-3 / args.steps
    e2e_s = max_over_ranks(torch, dist, world, max(wall / args.steps, ev_s))
    host_tables = float(np.mean([r["stages_ms"]["tables(host)"] for r in reports])) * 1e-3
    dev_s = max_over_ranks(torch, dist, world, float(np.mean([r["prove_ms"] for r in reports])) * 1e-3 - host_tables)
    stages = {k: float(np.mean([r["stages_ms"][k] for r in reports])) for k in reports[0]["stages_ms"]}
    kern = {k: v[0] / args.steps for k, v in sorted(prof.items(), key=lambda kv: -kv[1][0])}
    peak, peak_src = peaks()
    clocks = cs.summary()
    roof, roof_fft = rooflines(kern, FIB19, lmr, 1, clocks)
    h2d = sum(n * (1 << (lg - 4)) * 4 for lg, n, _ in FIB19)
    line = {"metric": "fib19.bf prove time", "value": dev_s, "unit": "s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dev_s * 1e3, "higher_is_better": False, "scaling": "strong", "vs_baseline": None,
            "dtype": "u32 (M31)", "data": "fib19.bf (reference example program), 199246 VM steps",
            "config": {"workload": "fib19_prove", "log_max_rows": lmr, "pcs": "pow 5, blowup 2x, 3 queries",
                       "twiddles": "tree of half_odds(26) cached per context (program-independent; recomputing it in every proof as the reference does adds 3.2 ms, profiles/r1_twiddle_cost.json); preprocessed tree recomputed every proof",
                       "columns": 213, "lde_cells": proof_lde_cells(FIB19, lmr), "l2": "working set (>20 GB) exceeds L2",
                       "parallelism": "single GPU (at --gpus N > 1 the same proof is split over N GPUs)",
                       "proof_sha256": __import__("hashlib").sha256(pr.json().encode()).hexdigest()},
            "e2e": {"value": e2e_s, "unit": "s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": proof_len,
                    "includes": "VM run, host table building, uploads, proof, proof readback"},
            "gpu_launches": int(launches // args.steps), "stages_ms": stages, "kernel_ms_per_proof": kern, "roofline": roof,
            "roofline_fft": roof_fft, "clocks": clocks, "verified": True, "e2e_preprocessed_cache": cached, "e2e_no_twiddle_cache": strict}
    if rank == 0 and not args.no_cpu_baseline:
        scaled, dt, thr, desc = cpu_prove_sample()
        line["cpu_baseline"] = {"value": scaled, "unit": "s", "cores": thr, "kind": "port", "sample": desc}
    if rank == 0:
        print(json.dumps(line))
    be.close()
    if world > 1:
        dist.destroy_process_group()


def bench_commit_sharded(args, torch, dist, rank, world, local, pkg, stream, be):
    """N > 1: ONE fib19 main-trace tree split over the ranks (strong scaling): column-sharded LDE, NCCL all-to-all to row
    ranges, row-sharded Merkle, sub-roots all-gathered (stwo-brainfuck_b200/sharded.py)."""
    sharded = importlib.import_module("stwo-brainfuck_b200.sharded")
    shape = [(max(4, lg - args.scale_down), n, k) for lg, n, k in FIB19]
    logs = [lg for lg, n, _ in shape for _ in range(n)]
    tw = be.precompute_twiddles(ROOT_LOG - args.scale_down)
    owner = sharded.assign_columns(logs, world)
    owned = {i: np.random.default_rng(0x5EED0000 + i).integers(0, P, size=1 << logs[i], dtype=np.uint32)
             for i in range(len(logs)) if owner[i] == rank}
    ops = sharded.CudaShardOps(pkg, be, tw, torch)
    h2d = sum(v.nbytes for v in owned.values())
    resident = {i: be.column(v) for i, v in owned.items()}

    def timed(cols, steps, warmup):
        for _ in range(warmup):
            root = sharded.sharded_commit(ops, dist, logs, cols, 1)
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            root = sharded.sharded_commit(ops, dist, logs, cols, 1)
        e1.record(stream)
        torch.cuda.synchronize()
        dist.barrier()
        return max_over_ranks(torch, dist, world, e0.elapsed_time(e1) / steps), root

    l0 = be.launch_count()
    with ClockSampler(local) as cs:
        ms, root = timed(resident, args.steps, args.warmup)           # inputs resident in HBM
        launches = be.launch_count() - l0
        ms_e2e, root2 = timed(owned, max(1, args.steps // 2), 1)      # host buffers: H2D inside the timed region
    assert (root == root2).all()
    roots = [None] * world
    dist.all_gather_object(roots, root.tolist())
    assert all(r == roots[0] for r in roots), "ranks disagree on the root"
    fft_b, merkle_b = commit_bytes(shape)
    alg = fft_b + merkle_b
    lde_bytes = sum(4 * (2 << lg) for lg in logs)
    line = {"metric": "LDE+commit throughput", "value": alg / (ms * 1e-3) / 1e9, "unit": "GB/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "u32 (M31)", "data": "synthetic",
            "config": {"workload": "fib19_commit", "columns": len(logs), "log_blowup": 1, "scale_down": args.scale_down,
                       "parallelism": f"column-sharded LDE -> all_to_all -> row-sharded Merkle x{world}",
                       "all_to_all_bytes_per_rank": int(lde_bytes * (world - 1) / world / world)},
            "e2e": {"value": alg / (ms_e2e * 1e-3) / 1e9, "unit": "GB/s", "ms_per_step": ms_e2e, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": 32},
            "gpu_launches": int(launches), "root": bytes(np.array(root, dtype=np.uint32)).hex(),
            "clocks": cs.summary()}
    if rank == 0:
        print(json.dumps(line))
    be.close()
    dist.destroy_process_group()


def bench_commit(args):
    torch, dist, rank, world, local, pkg, stream, be = setup(args)
    if world > 1:
        return bench_commit_sharded(args, torch, dist, rank, world, local, pkg, stream, be)
    shape = [(max(4, lg - args.scale_down), n, k) for lg, n, k in FIB19]
    tw = be.precompute_twiddles(ROOT_LOG - args.scale_down)
    logs = [lg for lg, n, _ in shape for _ in range(n)]
    host = [torch.from_numpy(np.random.default_rng(0x5EED0000 + i).integers(0, P, size=1 << lg, dtype=np.uint32).view(np.int32)).pin_memory()
            for i, lg in enumerate(logs)]   # same per-column seeds as the sharded arm: the roots must agree
    h2d = sum(t.numel() * 4 for t in host)
    resident = [be.column(t.numpy().view(np.uint32)) for t in host]

    def step_resident():
        cols = [c.clone() for c in resident]   # interpolate is in place; the clone is outside the algorithmic bytes
        be.interpolate_columns(cols, tw)
        ldes = be.evaluate_polynomials(cols, 1, tw)
        return be.merkle_commit(ldes)[1]

    def step_e2e():
        cols = [be.column(t.numpy().view(np.uint32)) for t in host]
        be.interpolate_columns(cols, tw)
        ldes = be.evaluate_polynomials(cols, 1, tw)
        return be.merkle_commit(ldes)[1]

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
        return max_over_ranks(torch, dist, world, e0.elapsed_time(e1) / steps), be.launch_count() - l0, root

    fft_b, merkle_b = commit_bytes(shape)
    alg = fft_b + merkle_b
    with ClockSampler(local) as cs:
        be.profile(True)
        be.profile_report()
        ms, launches, root = timed(step_resident, args.steps, args.warmup)
        prof = be.profile_report()
        be.profile(False)
        ms_e2e, _, root2 = timed(step_e2e, max(1, args.steps // 2), 1)
    assert (root == root2).all()
    kern = {k: v[0] / (args.steps + args.warmup) for k, v in prof.items()}  # scopes were recording during the warm-up too
    fft_ms = kern.get("fft_interpolate", 0) + kern.get("fft_evaluate", 0)
    peak, peak_src = peaks()
    roof = {"bound": "hbm", "kernel": "fft_kernel (interpolate + evaluate, all passes)",
            "achieved": fft_b / (fft_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s", "traffic": None, "peak_source": peak_src}
    roof["frac"] = roof["achieved"] / peak
    line = {"metric": "LDE+commit throughput", "value": world * alg / (ms * 1e-3) / 1e9, "unit": "GB/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u32 (M31)", "data": "synthetic",
            "config": {"workload": "fib19_commit", "columns": sum(n for _, n, _ in shape), "log_blowup": 1,
                       "scale_down": args.scale_down, "l2": "inputs (1 GB per step) exceed L2", "parallelism": f"replicas x{world}"},
            "e2e": {"value": world * alg / (ms_e2e * 1e-3) / 1e9, "unit": "GB/s", "ms_per_step": ms_e2e, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": 32},
            "gpu_launches": int(launches), "kernel_ms_per_step": kern, "roofline": roof, "clocks": cs.summary(),
            "root": bytes(np.array(root, dtype=np.uint32)).hex()}
    if rank == 0:
        print(json.dumps(line))
    be.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda")
    ap.add_argument("--workload", default="prove", choices=["prove", "commit"])
    ap.add_argument("--scale-down", type=int, default=0, help="commit workload only: shrink every column by 2^k rows")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--timeline", default=None, help="N > 1 prove workload: also dump every rank's profiling-scope timeline to <path>.rank<r>.json")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.workload == "commit":
        return bench_commit(args)
    return bench_prove(args)


if __name__ == "__main__":
    main()
