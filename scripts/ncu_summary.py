#!/usr/bin/env python
"""Summarise an .ncu-rep here (no GPU needed): headline metrics + per-phase instruction counts + top stalls.
usage: scripts/ncu_summary.py gpurun_out/prof.ncu-rep [kernel-index]"""
import collections, csv, io, subprocess, sys

rep = sys.argv[1]
which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
H = rows[0]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers',
        'launch__grid_size', 'launch__block_size', 'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
        'lts__t_sectors_srcunit_tex_op_red.sum', 'lts__t_sectors_srcunit_tex_op_red.sum.pct_of_peak_sustained_elapsed',
        'lts__d_atomic_input_cycles_active.avg.pct_of_peak_sustained_elapsed', 'lts__d_atomic_input_cycles_active.max.pct_of_peak_sustained_elapsed',
        'smsp__inst_executed_op_global_red.sum', 'l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_red.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'lts__t_requests_srcunit_tex_op_red.sum',
        'sm__inst_executed_pipe_lsu.sum', 'l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed']
want += [h for h in H if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('per_issue_active.ratio')]
r = rows[2 + which]
print('kernel:', r[H.index('Kernel Name')][:90])
for w in want:
  if w in H:
    v = r[H.index(w)]
    if w.startswith('smsp__average_warps_issue_stalled'):
      try:
        if float(v) < 0.15: continue
      except ValueError: pass
    print(f'  {w:82s} {v} {rows[1][H.index(w)]}')
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
blocks = [i for i, r in enumerate(rows) if r and r[0] == 'Kernel Name']
a = blocks[which]; b = blocks[which + 1] if which + 1 < len(blocks) else len(rows)
H = rows[a + 1]; data = [r for r in rows[a + 2:b] if len(r) == len(H)]
iI, iS, iSrc = H.index('Instructions Executed'), H.index('# Samples'), H.index('Source')
tot = sum(int(r[iI]) for r in data); tots = sum(int(r[iS]) for r in data)
print(f'total warp-instructions {tot/1e6:.1f}M, samples {tots}, sass lines {len(data)}')
def op(s):
  t = s.split()
  if t[0].startswith('@'): t = t[1:]
  return t[0].split('.')[0]
h = collections.Counter()
for r in data: h[op(r[iSrc])] += int(r[iI])
print('opcode mix (M):', [(k, round(v / 1e6, 1)) for k, v in h.most_common(22)])
stall_cols = [i for i, x in enumerate(H) if x.startswith('stall_') and 'Not Issued' not in x]
agg = collections.Counter()
for r in data:
  for i in stall_cols:
    if r[i] not in ('', '0'): agg[H[i]] += int(r[i])
print('stall samples:', [(k, v) for k, v in agg.most_common(10)])
print('top sampled instructions:')
for r in sorted(data, key=lambda r: -int(r[iS]))[:int(sys.argv[3]) if len(sys.argv) > 3 else 22]:
  st = sorted([(int(r[i]), H[i]) for i in stall_cols if r[i] not in ('', '0')], reverse=True)[:2]
  print(f"  {int(r[iS]):6d} {100*int(r[iS])/max(tots,1):5.1f}% inst={int(r[iI]):9d} {r[iSrc][:58]:58s} {st}")
