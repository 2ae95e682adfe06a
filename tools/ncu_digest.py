"""Digest an ncu report's exported CSV pages (developer tool).
usage: ncu_digest.py raw.csv sass.csv"""
import collections, csv, re, sys
raw, sass = sys.argv[1], sys.argv[2]
rows = list(csv.reader(open(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
keep = ['gpu__time_duration.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed', 'smsp__issue_active.avg.pct',
        'l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum ', 'l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum.per_second',
        'dram__bytes_read.sum ', 'dram__bytes_write.sum ', 'launch__registers_per_thread ', 'launch__shared_mem_per_block_dynamic',
        'sm__inst_executed.sum.per_cycle_elapsed', 'gpc__cycles_elapsed.max ', 'smsp__average_warps_issue_stalled', 'sm__warps_active.avg.pct',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum ', 'sm__throughput.avg.pct', 'l1tex__throughput.avg.pct',
        'lts__throughput.avg.pct', 'sm__inst_executed.sum ']
for h, u, v in zip(hdr, units, vals):
    if any(h.startswith(k.strip()) if k.endswith(' ') else h.startswith(k) for k in keep):
        if 'stalled' in h and float(v or 0) < 0.05: continue
        print(f'{h} [{u}] = {v}')
rows = list(csv.reader(open(sass)))
hdr = rows[1]; data = [r for r in rows[2:] if len(r) == len(hdr) and r[hdr.index('# Samples')].isdigit()]
iS = hdr.index('# Samples'); iSrc = hdr.index('Source'); iEx = hdr.index('Instructions Executed')
tot = sum(int(r[iS]) for r in data)
print('total samples', tot, 'static instrs', len(data), 'executed warp-instrs', sum(int(r[iEx]) for r in data))
byop = collections.Counter(); exop = collections.Counter()
for r in data:
    m = re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)', r[iSrc]); op = m.group(2).split('.')[0] if m else '?'
    byop[op] += int(r[iS]); exop[op] += int(r[iEx])
for op, c in byop.most_common(22): print(f'{op:12s} samples {c:8d} {100*c/tot:5.1f}%  executed {exop[op]}')
print('--- top instructions')
for r in sorted(data, key=lambda r: -int(r[iS]))[:int(sys.argv[3]) if len(sys.argv) > 3 else 30]: print(r[0][-5:], r[iS], r[iEx], r[iSrc][:100])
