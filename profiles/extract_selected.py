"""Write the metrics of a .ncu-rep that the round summaries quote as a small CSV (committed; the .ncu-rep
itself stays in gpurun_out/).  usage: python profiles/extract_selected.py <report.ncu-rep> <out.csv>"""
import csv
import subprocess
import sys

KEEP = ("gpu__time_duration", "dram__bytes", "launch__", "sm__inst_executed", "smsp__inst_executed.sum",
        "sm__icc_request_hit_rate", "sm__warps_active", "smsp__issue_active", "issue_stalled", "sm__throughput",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared", "smsp__inst_executed_op_shared", "sm__pipe_fma",
        "sm__pipe_fmaheavy", "smsp__inst_executed_pipe_lsu", "sm__cycles_elapsed.max", "lts__t_bytes.sum",
        "sm__sass_thread_inst_executed_op_ffma", "sm__inst_executed_pipe", "tensor", "tmem", "l1tex__data_pipe_lsu_wavefronts_mem_shared",
        "smsp__warp_issue_stalled", "sm__pipe_shared")
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h, u = rows[0], rows[1]
idx = [i for i, k in enumerate(h) if i < 3 or any(s in k for s in KEEP)]
with open(sys.argv[2], "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["metric", "unit"] + [f"launch{j}" for j in range(len(rows) - 2)])
    for i in idx:
        w.writerow([h[i], u[i]] + [r[i] for r in rows[2:]])
