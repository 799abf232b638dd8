#!/usr/bin/env python
"""Summarise an .ncu-rep (read on the CPU box): key metrics + opcode mix + hot regions.
usage: python tools/ncu_summary.py file.ncu-rep [--regions]"""
import csv, io, subprocess, sys
from collections import Counter

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
keys = ['gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__waves_per_multiprocessor',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__inst_executed.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__warps_eligible.avg.per_cycle_active', 'smsp__warps_active.avg.per_cycle_active', 'sm__cycles_elapsed.avg',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'smsp__cycles_active.avg',
        'sm__cycles_active.avg', 'smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio']
print("== metrics")
for i, h in enumerate(hdr):
    if h in keys:
        print("%-70s %-16s %s" % (h, units[i], vals[i]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr, data = rows[1], rows[2:]
iS, iE, iT, iSm = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Avg. Threads Executed"), hdr.index("# Samples")
tot = sum(int(r[iE]) for r in data)
print("== total warp instructions %.3f G, SASS lines %d" % (tot / 1e9, len(data)))
c = Counter()
for r in data:
    parts = r[iS].split()
    op = parts[1] if parts[0].startswith('@') else parts[0]
    c[op.split('.')[0]] += int(r[iE])
print("== opcode mix")
print("  ".join("%s %.1f%%" % (k, v / tot * 100) for k, v in c.most_common(24)))
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
st = Counter()
for r in data:
    for i in stall_cols:
        try:
            st[hdr[i]] += int(r[i])
        except ValueError:
            pass
ts = sum(st.values()) or 1
print("== stall samples")
print("  ".join("%s %.1f%%" % (k, v / ts * 100) for k, v in st.most_common(10)))
if "--regions" in sys.argv:
    segs = []
    for i, r in enumerate(data):
        e = int(r[iE])
        if segs and segs[-1][3] > 0 and abs(e - segs[-1][3]) <= 0.15 * max(e, segs[-1][3]):
            s = segs[-1]; s[1] = i; s[2] += e; s[4] += int(r[iSm])
        else:
            segs.append([i, i, e, e, int(r[iSm])])
    print("== regions (>0.5% of instructions)")
    for s in segs:
        if s[2] > 0.005 * tot:
            print("[%4d-%4d] n=%3d per-instr=%8.1fM total=%5.2f%% samples=%d" % (s[0], s[1], s[1] - s[0] + 1, s[3] / 1e6, s[2] / tot * 100, s[4]))
if "--dump" in sys.argv:
    lo, hi = int(sys.argv[sys.argv.index("--dump") + 1]), int(sys.argv[sys.argv.index("--dump") + 2])
    for i in range(lo, hi + 1):
        r = data[i]
        print("%5d %9.1fM thr=%5s smp=%6s  %s" % (i, int(r[iE]) / 1e6, r[iT], r[iSm], r[iS].strip()))
