#!/usr/bin/env python
"""Summarise an .ncu-rep (read on the CPU box): key metrics of every captured launch + top stall sites.
usage: tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/<name>.txt"""
import csv, io, subprocess, sys, collections
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
KEYS = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.avg.per_second"]
print(f"# ncu summary of {rep.split('/')[-1]} (captured with --set full --clock-control none; times are serialised, cold-cache)")
for r in rows[2:]:
    if len(r) != len(hdr):
        continue
    print("-" * 100)
    for i, h in enumerate(hdr):
        if h in KEYS:
            print(f"{h:86s} {units[i]:16s} {r[i]}")
    for i, h in enumerate(hdr):
        if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
            try:
                v = float(r[i])
            except ValueError:
                continue
            if v >= 0.05:
                print(f"  stall {h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]:28s} {v:.3f} per issue")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
if len(rows) > 3 and "Address" in rows[1]:
    h = rows[1]; data = [r for r in rows[2:] if len(r) == len(h)]
    ia, isrc, iex, ism = h.index("Address"), h.index("Source"), h.index("Instructions Executed"), h.index("Warp Stall Sampling (All Samples)")
    tot = sum(int(r[iex]) for r in data) or 1; tots = sum(int(r[ism]) for r in data) or 1
    base = int(data[0][ia], 16)
    print("-" * 100)
    print(f"SASS instructions: {len(data)}; warp instructions executed: {tot}; stall samples: {tots}")
    ops = collections.Counter()
    for r in data:
        t = r[isrc].split()
        op = (t[1] if t and t[0].startswith('@') else t[0]).split('.')[0] if t else '?'
        ops[op] += int(r[iex])
    print("executed opcode mix: " + ", ".join(f"{o} {100*c/tot:.1f}%" for o, c in ops.most_common(12)))
    print("top stall sites (offset, % of samples, executions, SASS):")
    for r in sorted(data, key=lambda r: -int(r[ism]))[:12]:
        print(f"  {int(r[ia],16)-base:#07x} {100*int(r[ism])/tots:5.1f}%  {r[iex]:>10s}  {r[isrc][:84]}")
