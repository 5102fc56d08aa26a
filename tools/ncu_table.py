#!/usr/bin/env python
"""Prints the judged metrics of every launch in an .ncu-rep (needs ncu on PATH; no GPU): python tools/ncu_table.py file.ncu-rep [--csv out.csv]"""
import csv, io, subprocess, sys
M = [("ms", "gpu__time_duration.sum"), ("inst", "smsp__inst_executed.sum"), ("alu%", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
     ("fma%", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"), ("lsu%", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
     ("issue%", "smsp__issue_active.avg.pct_of_peak_sustained_active"), ("warps%", "sm__warps_active.avg.pct_of_peak_sustained_active"),
     ("regs", "launch__registers_per_thread"), ("dram_rd_GB", "dram__bytes_read.sum"), ("dram_wr_GB", "dram__bytes_write.sum"),
     ("dram%", "dram__throughput.avg.pct_of_peak_sustained_elapsed"), ("l1%", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed"),
     ("smem_conf", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"), ("smem_wave", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"),
     ("local_ld", "smsp__inst_executed_op_local_ld.sum"), ("local_st", "smsp__inst_executed_op_local_st.sum"),
     ("st_math", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"),
     ("st_long", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"),
     ("st_short", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio"),
     ("st_wait", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"),
     ("st_bar", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"),
     ("st_notsel", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"),
     ("st_disp", "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio"),
     ("st_lg", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio"),
     ("st_mio", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio"),
     ("st_imc", "smsp__average_warps_issue_stalled_imc_miss_per_issue_active.ratio"),
     ("st_nninst", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio")]
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
out = []
for r in rows[2:]:
    d = dict(zip(hdr, r))
    name = d["Kernel Name"].replace("void ", "").split("(")[0]
    rec = {"kernel": name, "grid": d["Grid Size"], "block": d["Block Size"]}
    for k, m in M:
        v = d.get(m, "")
        if k == "ms" and v:
            u = units[hdr.index(m)]
            v = float(v) * {"us": 1e-3, "ns": 1e-6, "ms": 1.0, "s": 1e3}.get(u, 1.0)
        try:
            rec[k] = round(float(v), 3)
        except ValueError:
            rec[k] = v
    out.append(rec)
keys = ["kernel", "grid", "block"] + [k for k, _ in M]
if "--csv" in sys.argv:
    w = csv.DictWriter(open(sys.argv[sys.argv.index("--csv") + 1], "w", newline=""), fieldnames=keys)
    w.writeheader(); w.writerows(out)
for rec in out:
    print(rec["kernel"], rec["grid"], rec["block"])
    print("   " + "  ".join(f"{k}={rec[k]}" for k, _ in M))
