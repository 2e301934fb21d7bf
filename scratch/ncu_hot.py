"""Summarise an ncu report: key raw metrics per kernel + hottest SASS lines with stall reasons.
usage: python scratch/ncu_hot.py report.ncu-rep [kernel-regex] [topN]"""
import csv, subprocess, sys, io, collections
rep = sys.argv[1]; kre = sys.argv[2] if len(sys.argv) > 2 else None; top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw))); hdr, units, data = rows[0], rows[1], rows[2:]
want = ['Kernel Name', 'gpu__time_duration.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed', 'dram__bytes_read.sum ', 'dram__bytes_write.sum ',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct', 'lts__t_sectors.avg.pct', 'sm__cycles_elapsed.avg.per_second',
        'smsp__issue_active.avg.pct', 'launch__registers_per_thread ', 'sm__inst_executed.sum ', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum ']
for i, h in enumerate(hdr):
    if any((h + ' ').startswith(w) or w.strip() == h for w in want) or h in ('dram__bytes_read.sum', 'dram__bytes_write.sum'):
        print(h, units[i], [d[i][:40] for d in data])
cmd = ["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name-base", "mangled"] + (["-k", "regex:" + kre] if kre else [])
src = subprocess.run(cmd, capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = next(r for r in rows if r and r[0] == 'Address'); ix = {h: i for i, h in enumerate(hdr)}
data = []; seen = set()
for r in rows:
    if r and r[0].startswith('0x'):
        if r[0] in seen: break
        seen.add(r[0]); data.append(r)
tot = sum(int(r[ix['# Samples']]) for r in data); print('total samples', tot, 'instrs', len(data))
stalls = [h for h in hdr if h.startswith('stall_') and 'Not' not in h]
agg = collections.Counter()
for r in data:
    for h in stalls: agg[h[6:]] += int(r[ix[h]])
print('stall totals', agg.most_common(8))
for r in sorted(data, key=lambda r: -int(r[ix['# Samples']]))[:top]:
    s = {h[6:]: int(r[ix[h]]) for h in stalls if int(r[ix[h]]) > 0}
    s = dict(sorted(s.items(), key=lambda kv: -kv[1])[:3])
    print(data.index(r), r[ix['# Samples']], r[ix['Instructions Executed']], r[1].strip()[:80], s)
