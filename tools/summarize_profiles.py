#!/usr/bin/env python
"""Turns the raw ncu CSVs of tools/profile_r2.sh (gpurun_out/) into the tracked summaries under profiles/:
  r2_launches_bench_prove_summary.csv   per-kernel share of the serialised launch list of the bench command
  r2_merkle_traffic.json                DRAM bytes of every Merkle launch of one warm fib19 proof, summed (bench.py reads it)
  r2_ncu_full_<name>.csv                the judged metrics of every launch in each --set full capture"""
import collections
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, PR = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")


def read_ncu_csv(path):
    rows = list(csv.reader(open(path)))
    h = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr = rows[h]
    return [dict(zip(hdr, r)) for r in rows[h + 1:] if len(r) == len(hdr)]


def short(name):
    m = re.match(r"(?:void )?(?:sb::)?([A-Za-z0-9_]+)(<[^(]*>)?", name)
    return (m.group(1) + (m.group(2) or "")) if m else name


def launches():
    src = os.path.join(G, "r2_launches_bench_prove.csv")
    if not os.path.exists(src):
        return
    rows = [r for r in read_ncu_csv(src) if r["Metric Name"] == "gpu__time_duration.sum"]
    unit = {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "nsecond": 1e-6, "ms": 1.0, "msecond": 1.0}
    tot = collections.defaultdict(lambda: [0.0, 0])
    for r in rows:
        ms = float(r["Metric Value"].replace(",", "")) * unit.get(r["Metric Unit"], 1e-6)
        k = short(r["Kernel Name"])
        tot[k][0] += ms
        tot[k][1] += 1
    allms = sum(v[0] for v in tot.values())
    with open(os.path.join(PR, "r2_launches_bench_prove_summary.csv"), "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none over `python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extra`\n")
        f.write("# (every proof of the run, the legs after the headline included; per-launch times are serialised and cold-cache: use the SHARES)\n")
        f.write("kernel,launches,total_ms,share\n")
        for k, (ms, n) in sorted(tot.items(), key=lambda kv: -kv[1][0]):
            f.write(f"\"{k}\",{n},{ms:.3f},{ms / allms:.4f}\n")
    merkle = sum(v[0] for k, v in tot.items() if k.startswith("commit_"))
    print("launch list: %d launches, %.1f ms, Merkle share %.3f" % (len(rows), allms, merkle / allms))


def merkle_traffic():
    src = os.path.join(G, "r2_merkle_traffic.csv")
    if not os.path.exists(src):
        return
    rows = read_ncu_csv(src)
    per = collections.OrderedDict()
    for r in rows:
        per.setdefault(r["ID"], {"kernel": short(r["Kernel Name"]), "grid": r["Grid Size"]})[r["Metric Name"]] = float(r["Metric Value"].replace(",", ""))
    ids = list(per)
    counts = json.loads(open(os.path.join(G, "r2_merkle_traffic.out")).read().strip().splitlines()[-1])["launches_per_proof"]
    half = len(ids) // 2                     # two proofs, the same Merkle launches in each: the second one is warm
    second = [per[i] for i in ids[half:]]
    rd = sum(x.get("dram__bytes_read.sum", 0) for x in second)
    wr = sum(x.get("dram__bytes_write.sum", 0) for x in second)
    sys.path.insert(0, ROOT)
    import bench
    alg, comp, done = bench.proof_merkle_stats(bench.FIB19, 24)
    out = {"program": "fib19.bf", "log_max_rows": 24, "merkle_launches_per_proof": len(second), "kernel_launches_per_proof": counts[-1],
           "dram_bytes_read": rd, "dram_bytes_written": wr, "dram_bytes_per_proof": rd + wr, "algorithmic_bytes": alg,
           "traffic_over_algorithmic": (rd + wr) / alg,
           "note": "ncu dram__bytes_read.sum + dram__bytes_write.sum summed over every commit_* launch of the second of two fib19 proofs "
                   "(profiles/r2_merkle_traffic.csv.gz); algorithmic = R*(4C + 32 + 64*[children]) per layer. The sub-tree kernel hands "
                   "children over in shared memory, so layers 2^19..2^10 are not re-read; the lane-repeated main tree reads 1/16 of its leaves."}
    json.dump(out, open(os.path.join(PR, "r2_merkle_traffic.json"), "w"), indent=1)
    import gzip
    import shutil
    with open(src, "rb") as a, gzip.open(os.path.join(PR, "r2_merkle_traffic.csv.gz"), "wb") as b:
        shutil.copyfileobj(a, b)
    print("merkle traffic: %d launches, %.2f GB read + %.2f GB written = %.3f x algorithmic" % (len(second), rd / 1e9, wr / 1e9, (rd + wr) / alg))


KEEP = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "smsp__average_warp_latency_issue_stalled_math_pipe_throttle.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"]


def full_reports():
    """r2_ncu_full_<name>.raw.csv (the raw page of a --set full capture, written on the GPU box by tools/profile_r2.sh) ->
    profiles/r2_ncu_full_<name>.csv with the judged metrics of every captured launch."""
    for rep in sorted(os.listdir(G)):
        if not (rep.startswith("r2_ncu_full_") and rep.endswith(".raw.csv")):
            continue
        rows = list(csv.reader(open(os.path.join(G, rep))))
        if len(rows) < 3:
            continue
        hdr = rows[0]
        idx = {n: i for i, n in enumerate(hdr)}
        cols = [c for c in KEEP if c in idx]
        out = os.path.join(PR, rep.replace(".raw.csv", ".csv"))
        with open(out, "w") as f:
            f.write("# ncu --set full --clock-control none; one row per captured launch\n")
            f.write("# units: " + ", ".join("%s [%s]" % (c, rows[1][idx[c]]) for c in cols if rows[1][idx[c]]) + "\n")
            f.write(",".join(["kernel", "grid", "block"] + cols) + "\n")
            for r in rows[2:]:
                if len(r) != len(hdr):
                    continue
                f.write(",".join(['"%s"' % short(r[idx["Kernel Name"]]), '"%s"' % r[idx["Grid Size"]], '"%s"' % r[idx["Block Size"]]] + [r[idx[c]].replace(",", "") for c in cols]) + "\n")
        print("wrote", out)


if __name__ == "__main__":
    launches()
    merkle_traffic()
    full_reports()
    for f in os.listdir(G):
        if f.startswith("r2_sanitizer_"):
            import shutil
            shutil.copy(os.path.join(G, f), os.path.join(PR, f))
