#!/usr/bin/env python
"""Summarise an .ncu-rep: key raw metrics per kernel + the hottest SASS lines (stall samples).
usage: python tools_ncu_summary.py report.ncu-rep [kernel-substring] [min_fraction]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; kern = sys.argv[2] if len(sys.argv) > 2 else None; frac = float(sys.argv[3]) if len(sys.argv) > 3 else 0.01
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ['Kernel Name','gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','sm__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__throughput.avg.pct_of_peak_sustained_elapsed','lts__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__t_sector_hit_rate.pct','lts__t_sector_hit_rate.pct','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','launch__occupancy_limit_registers','launch__occupancy_limit_shared_mem','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__inst_executed.sum','sm__inst_executed_pipe_alu.sum','sm__inst_executed_pipe_lsu.sum','sm__inst_executed_pipe_xu.sum','sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active','sm__pipe_xu_cycles_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_uniform.sum','smsp__thread_inst_executed_per_inst_executed.ratio','smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio','smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_wait_per_issue_active.ratio','smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio','smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio','smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio','smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio','smsp__average_warps_issue_stalled_membar_per_issue_active.ratio']
idx = {h: i for i, h in enumerate(hdr)}
for r in rows[2:]:
    if kern and kern not in r[idx['Kernel Name']]: continue
    print('=====')
    for w in want:
        if w in idx: print(f"{w} = {r[idx[w]]} {units[idx[w]]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"] + (["--kernel-name", "regex:" + kern] if kern else []), capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = next(i for i, r in enumerate(rows) if 'Source' in r and 'Instructions Executed' in r)
hd = rows[h]; si = hd.index('Source'); a = hd.index('Warp Stall Sampling (All Samples)'); e = hd.index('Instructions Executed')
data = []
for r in rows[h+1:]:
    try: data.append((int(r[a] or 0), int(r[e] or 0), r[si]))
    except Exception: pass
tot = sum(d[0] for d in data) or 1
print('total samples', tot, 'sass lines', len(data), 'total warp-instr', sum(d[1] for d in data))
for j, d in enumerate(data):
    if d[0] > tot * frac: print(j, f"{100*d[0]/tot:5.1f}%", d[1], d[2][:100])
