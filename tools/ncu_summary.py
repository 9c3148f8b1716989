#!/usr/bin/env python3
"""Summarise an .ncu-rep: headline metrics per kernel + per-phase (BAR-delimited) stall / opcode mix from the source page.
usage: tools/ncu_summary.py gpurun_out/prof.ncu-rep [kernel-regex]"""
import collections, csv, io, subprocess, sys
rep = sys.argv[1]; kre = sys.argv[2] if len(sys.argv) > 2 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'smsp__inst_executed.sum',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'sm__cycles_elapsed.max', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'launch__occupancy_limit_warps', 'sm__maximum_warps_per_active_cycle_pct', 'launch__grid_size', 'launch__block_size']
idx = [(w, hdr.index(w)) for w in want if w in hdr]
for r in rows[2:]:
    print('---')
    for w, i in idx:
        print('  %-72s %s %s' % (w, r[i], units[i]))
args = ["ncu", "-i", rep, "--page", "source", "--csv"] + (["--kernel-name", "regex:" + kre] if kre else [])
src = subprocess.run(args, capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
secs = [i for i, r in enumerate(rows) if r and r[0] == 'Kernel Name'] + [len(rows)]
for a, b in zip(secs[:-1], secs[1:]):
    print('=== source page:', rows[a][1][:60])
    hdr = rows[a + 1]; data = [r for r in rows[a + 2:b] if len(r) == len(hdr)]
    ci = {h: i for i, h in enumerate(hdr)}
    stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
    phase = 0; per = collections.defaultdict(collections.Counter); inst = collections.Counter(); samp = collections.Counter(); opc = collections.Counter()
    for r in data:
        s = r[ci['Source']]
        if 'BAR.SYNC' in s: phase += 1
        n = int(r[ci['Instructions Executed']] or 0); sm = int(r[ci['# Samples']] or 0)
        inst[phase] += n; samp[phase] += sm
        t = s.split(); op = t[1] if t[0].startswith('@') else t[0]
        opc[(phase, op.split('.')[0])] += n
        for st in stalls: per[phase][st] += int(r[ci[st]] or 0)
    tots = sum(samp.values()) or 1
    for ph in sorted(inst):
        t = sum(per[ph].values()) or 1
        print(' phase %d: inst %d (%.0f%%) samples %d (%.0f%%)' % (ph, inst[ph], 100 * inst[ph] / (sum(inst.values()) or 1), samp[ph], 100 * samp[ph] / tots))
        print('    stalls:', ', '.join('%s %.0f%%' % (k[6:], 100 * v / t) for k, v in per[ph].most_common(7)))
        ops = sorted([(k[1], v) for k, v in opc.items() if k[0] == ph], key=lambda kv: -kv[1])
        print('    ops:', ', '.join('%s %d' % kv for kv in ops[:14]))
