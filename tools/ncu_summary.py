"""Summarises an exported ncu report (raw + source CSV pages) for profiles/."""
import csv, collections, sys
raw, src, steps_per_launch = sys.argv[1], sys.argv[2], float(sys.argv[3])
rows = list(csv.reader(open(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__block_size', 'sm__warps_active.avg.per_cycle_active', 'smsp__inst_executed.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'sm__inst_issued.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fp32.avg.pct_of_peak_sustained_active',
        'sm__cycles_elapsed.avg', 'sm__cycles_elapsed.avg.per_second', 'sm__throughput.avg.pct_of_peak_sustained_elapsed']
for w in want:
    if w in hdr:
        i = hdr.index(w); print(f"{w:82s} {units[i]:16s} {vals[i]}")
for i, h in enumerate(hdr):
    if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('per_issue_active.ratio') and float(vals[i]) > 0.05:
        print(f"stall {h.replace('smsp__average_warps_issue_stalled_','').replace('_per_issue_active.ratio',''):30s} {vals[i]}")
srows = list(csv.reader(open(src)))
h = srows[1]; data = srows[2:]
iS, iE, iSamp = h.index('Source'), h.index('Instructions Executed'), h.index('# Samples')
wsteps = steps_per_launch / 32
by, bys = collections.Counter(), collections.Counter()
for r in data:
    op = [o for o in r[iS].split() if not o.startswith('@')][0].split('.')[0]
    by[op] += int(r[iE]); bys[op] += int(r[iSamp])
ts, tot = sum(bys.values()), sum(by.values())
print(f"\nwarp-instructions executed per warp-step (32 ray-steps): {tot / wsteps:.1f}")
print("opcode      inst/warp-step   share of stall samples")
for op, c in by.most_common(24):
    print(f"{op:10s} {c / wsteps:10.2f}        {bys[op] / ts * 100:5.1f}%")
