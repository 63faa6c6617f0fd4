#!/usr/bin/env python
"""Summarise an .ncu-rep (ncu -i X --page raw --csv) into a markdown table of the metrics the
roofline discussion uses, and print per-launch DRAM traffic as JSON on the last line.
usage: tools/ncu_summary.py <report.ncu-rep> [title]"""
import csv, io, json, subprocess, sys
rep = sys.argv[1]
title = sys.argv[2] if len(sys.argv) > 2 else rep
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units, body = rows[0], rows[1], rows[2:]
want = [("gpu__time_duration.sum", "duration"), ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
        ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM % of peak"),
        ("lts__t_sector_hit_rate.pct", "L2 hit %"), ("smsp__inst_executed.sum", "warp instructions"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
        ("launch__registers_per_thread", "registers/thread"), ("launch__shared_mem_per_block_dynamic", "dyn smem/block"),
        ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts")]
ci = {h: i for i, h in enumerate(hdr)}
names = [r[ci["Kernel Name"]].split("(")[0] for r in body]
print(f"### {title}\n")
print("| metric | " + " | ".join(names) + " |")
print("|---|" + "---|" * len(names))
traffic = {}
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
for key, label in want:
    if key not in ci:
        continue
    i = ci[key]
    print(f"| {label} ({units[i]}) | " + " | ".join(f"{float(r[i]):,.3f}".rstrip("0").rstrip(".") if r[i] else "" for r in body) + " |")
for r, n in zip(body, names):
    rd = float(r[ci["dram__bytes_read.sum"]]) * scale.get(units[ci["dram__bytes_read.sum"]], 1)
    wr = float(r[ci["dram__bytes_write.sum"]]) * scale.get(units[ci["dram__bytes_write.sum"]], 1)
    traffic[n] = {"dram_bytes": rd + wr, "ms": float(r[ci["gpu__time_duration.sum"]]) * (1e-3 if units[ci["gpu__time_duration.sum"]] == "us" else 1.0)}
print()
print("TRAFFIC " + json.dumps(traffic))
