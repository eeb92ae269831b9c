"""Summarise an ncu report (raw + source pages) of the tile kernel: key metrics, opcode mix, stall reasons, hot SASS.
Usage: python scripts/ncu_summary.py gpurun_out/x/prof_tile.ncu-rep [n_hot]"""
import csv, io, subprocess, sys
from collections import Counter

rep = sys.argv[1]
nhot = int(sys.argv[2]) if len(sys.argv) > 2 else 16
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
keys = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__block_size', 'launch__shared_mem_per_block_dynamic', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__cycles_elapsed.max',
        'sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_sector_pipe_lsu_mem_global_op_ld_hit_rate.pct',
        'lts__t_sectors_srcunit_tex_op_read.sum', 'lts__throughput.avg.pct_of_peak_sustained_elapsed']
for r in rows[2:3]:
    for k in keys:
        if k in hdr:
            i = hdr.index(k)
            print(f"{k} = {r[i]} {units[i]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'Address'][0]
hdr = rows[hi]; ix = {k: i for i, k in enumerate(hdr)}
out = []
for r in rows[hi + 1:]:
    if r and r[0] == 'Kernel Name':
        break
    if len(r) >= len(hdr):
        out.append(r)
ti = sum(int(r[ix['Instructions Executed']]) for r in out); ts = sum(int(r[ix['# Samples']]) for r in out)
print(f"-- SASS: {len(out)} instructions, executed {ti}, samples {ts}")
c, s = Counter(), Counter()
for r in out:
    op = [t for t in r[ix['Source']].split() if not t.startswith('@')][0].split('.')[0]
    c[op] += int(r[ix['Instructions Executed']]); s[op] += int(r[ix['# Samples']])
for op, n in c.most_common(14):
    print(f"  {op:10s} {n:12d} {100*n/ti:5.1f}%  samples {100*s[op]/ts:5.1f}%")
print("-- stall reasons (share of samples)")
for k in [k for k in hdr if k.startswith('stall_') and 'Not Issued' not in k]:
    v = sum(int(r[ix[k]]) for r in out)
    if v > ts * 0.015:
        print(f"  {k:26s} {100*v/ts:5.1f}%")
print("-- hottest SASS")
for r in sorted(out, key=lambda r: -int(r[ix['# Samples']]))[:nhot]:
    print(f"  {int(r[ix['# Samples']]):7d} {100*int(r[ix['# Samples']])/ts:5.1f}%  {r[ix['Source']][:90]}")
