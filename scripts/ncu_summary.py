#!/usr/bin/env python3
"""Summarise an .ncu-rep (ncu --set full) into a markdown table: usage ncu_summary.py rep.ncu-rep out.md [title]"""
import csv, subprocess, sys, io
rep, out = sys.argv[1], sys.argv[2]
title = sys.argv[3] if len(sys.argv) > 3 else rep
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
want = [("gpu__time_duration.sum", "duration"), ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
        ("lts__t_bytes.sum", "L2 bytes"), ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
        ("launch__registers_per_thread", "registers/thread"), ("launch__shared_mem_per_block_dynamic", "dyn smem/block"),
        ("launch__shared_mem_per_block_static", "static smem/block"),
        ("launch__occupancy_limit_registers", "occupancy limit (regs) blocks"), ("launch__occupancy_limit_shared_mem", "occupancy limit (smem) blocks"),
        ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
        ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier"),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long scoreboard"),
        ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short scoreboard"),
        ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "stall MIO throttle"),
        ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall LG throttle")]
lines = [f"# {title}", "", f"source: `ncu --set full --clock-control none --import-source on` (`{rep.split('/')[-1]}`, not committed: binary)", ""]
for r in rows[2:]:
    name = r[idx["Kernel Name"]]
    lines.append(f"## `{name[:110]}`")
    lines.append("")
    lines.append("| metric | value | unit |")
    lines.append("|---|---:|---|")
    for key, label in want:
        if key in idx and r[idx[key]] != "":
            lines.append(f"| {label} (`{key}`) | {r[idx[key]]} | {units[idx[key]]} |")
    try:
        rd = float(r[idx["dram__bytes_read.sum"]].replace(",", "")); wr = float(r[idx["dram__bytes_write.sum"]].replace(",", ""))
        lines.append(f"| **traffic = read + write** | {rd + wr:.1f} | {units[idx['dram__bytes_read.sum']]} |")
    except Exception:
        pass
    lines.append("")
open(out, "w").write("\n".join(lines) + "\n")
print("\n".join(lines))
