#!/usr/bin/env python3
"""Summarise an .ncu-rep here (no GPU): key raw metrics + stall samples bucketed over the SASS, by source file:line.
usage: tools/ncu_summary.py gpurun_out/prof_x.ncu-rep [out.csv]"""
import csv, subprocess, sys, io, collections
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__registers_per_thread',
        'launch__grid_size', 'launch__block_size', 'launch__shared_mem_per_block_dynamic', 'launch__shared_mem_per_block_static',
        'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__sass_thread_inst_executed_op_dfma_pred_on.sum', 'smsp__sass_thread_inst_executed_op_dadd_pred_on.sum',
        'smsp__sass_thread_inst_executed_op_dmul_pred_on.sum', 'l1tex__t_sector_pipe_lsu_mem_local_op_ld_hit_rate.pct',
        'sass__inst_executed_local_loads', 'sass__inst_executed_local_stores', 'sass__inst_executed_shared_loads', 'sass__inst_executed_shared_stores',
        'smsp__pcsamp_sample_count', 'smsp__inst_executed.sum', 'lts__t_sector_hit_rate.pct', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum']
out = []
for h, u, v in zip(hdr, units, vals):
    if h in want or h.startswith('smsp__pcsamp_warps_issue_stalled') and not h.endswith('not_issued'):
        out.append((h, u, v))
for h, u, v in out:
    print("%-75s %-10s %s" % (h, u, v))
if len(sys.argv) > 2:
    with open(sys.argv[2], 'w') as f:
        f.write('metric,unit,value\n')
        for h, u, v in out:
            f.write('%s,%s,%s\n' % (h, u, v))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
tot = sum(int(r[ix['# Samples']]) for r in data) or 1
print("total samples", tot, "SASS instructions", len(data))
B = max(500, len(data) // 40)
for b in range(0, len(data), B):
    seg = data[b:b + B]
    s = sum(int(r[ix['# Samples']]) for r in seg)
    def col(n): return sum(int(r[ix[n]]) for r in seg)
    print("%6d  %5.1f%%  no_inst %4.1f%% long_sb %4.1f%% short_sb %4.1f%% wait %4.1f%% barrier %4.1f%% math %4.1f%% sel %4.1f%% exec %11d" % (
        b, 100 * s / tot, 100 * col('stall_no_inst') / tot, 100 * col('stall_long_sb') / tot, 100 * col('stall_short_sb') / tot, 100 * col('stall_wait') / tot,
        100 * col('stall_barrier') / tot, 100 * col('stall_math') / tot, 100 * col('stall_selected') / tot, col('Instructions Executed')))
