"""Summarises an exported ncu report for profiles/:
    ncu -i rep.ncu-rep --page raw --csv > raw.csv;  ncu -i rep.ncu-rep --page source --csv > src.csv
    python tools/ncu_summary.py raw.csv src.csv <launch index in the report> <ray-steps of that launch>
Raw page: one row per captured launch.  Source page: one section per launch ("Kernel Name" line, header line, one row per
SASS instruction with its executed count and stall samples)."""
import collections
import csv
import sys

raw, src, launch, steps_per_launch = sys.argv[1], sys.argv[2], int(sys.argv[3]), float(sys.argv[4])
rows = list(csv.reader(open(raw)))
hdr, units, vals = rows[0], rows[1], rows[2 + launch]
want = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__block_size', 'sm__warps_active.avg.per_cycle_active', 'smsp__inst_executed.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'sm__inst_issued.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
        'l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fp32.avg.pct_of_peak_sustained_active',
        'sm__cycles_elapsed.avg', 'sm__cycles_elapsed.avg.per_second', 'sm__throughput.avg.pct_of_peak_sustained_elapsed']
for w in want:
    if w in hdr:
        i = hdr.index(w)
        print(f"{w:82s} {units[i]:16s} {vals[i]}")
for i, h in enumerate(hdr):
    if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('per_issue_active.ratio'):
        try:
            if float(vals[i]) > 0.05:
                print(f"stall {h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''):30s} {vals[i]}")
        except ValueError:
            pass
sections, cur = [], None
for r in csv.reader(open(src)):
    if r and r[0] == 'Kernel Name':
        cur = {'name': r[1] if len(r) > 1 else '', 'hdr': None, 'rows': []}
        sections.append(cur)
    elif cur is not None and cur['hdr'] is None:
        cur['hdr'] = r
    elif cur is not None and r:
        cur['rows'].append(r)
# every launch appears twice in the source page (SASS view and source view list the same instructions): keep one of each pair
if len(sections) >= 2 * (len(rows) - 2) - 1:
    sections = sections[::2]
sec = sections[launch]
h = sec['hdr']
iS, iE, iSamp = h.index('Source'), h.index('Instructions Executed'), h.index('# Samples')
wsteps = steps_per_launch / 32
by, bys = collections.Counter(), collections.Counter()
for r in sec['rows']:
    if len(r) <= max(iS, iE, iSamp) or not r[iS].strip():
        continue
    op = [o for o in r[iS].split() if not o.startswith('@')][0].split('.')[0]
    by[op] += int(r[iE]); bys[op] += int(r[iSamp])
ts, tot = sum(bys.values()), sum(by.values())
fp64 = sum(c for op, c in by.items() if op in ('DFMA', 'DMUL', 'DADD', 'DSETP'))
print(f"\nsource section: {sec['name'][:110]}")
print(f"warp-instructions executed per warp-step (32 ray-steps): {tot / wsteps:.1f}   (fp64-pipe: {fp64 / wsteps:.1f})")
print("opcode      inst/warp-step   share of stall samples")
for op, c in by.most_common(26):
    print(f"{op:10s} {c / wsteps:10.2f}        {bys[op] / ts * 100:5.1f}%")
