"""Turns ncu outputs under gpurun_out/ into the small text summaries committed under profiles/."""
import collections, csv, re, sys

def launch_summary(path, out):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    idx = {h: i for i, h in enumerate(rows[0])}
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[1:]:
        if r[idx['Metric Name']] != 'gpu__time_duration.sum':
            continue
        name = r[idx['Kernel Name']]
        m = re.search(r'k_tc<cffm::tc::(\w+)<?([^>]*)', name)
        short = ('k_tc<' + m.group(1) + '<' + m.group(2) + '>>') if m else name.split('(')[0][:70]
        v = float(r[idx['Metric Value']].replace(',', ''))
        v *= {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 's': 1e3}.get(r[idx['Metric Unit']], 1e-6)
        agg[short][0] += 1; agg[short][1] += v
    tot = sum(v[1] for v in agg.values())
    with open(out, 'w') as f:
        f.write('# ncu --metrics gpu__time_duration.sum --clock-control none: per-kernel device time (cold, serialised)\n')
        f.write('# source: %s ; total %.3f ms over %d launches\n' % (path, tot, sum(v[0] for v in agg.values())))
        f.write('%-64s %6s %11s %7s %9s\n' % ('kernel', 'n', 'total_ms', 'share', 'avg_ms'))
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write('%-64s %6d %11.3f %7.4f %9.4f\n' % (k, v[0], v[1], v[1] / tot, v[1] / v[0]))

def full_summary(path, out):
    rows = list(csv.reader(open(path)))
    hdr = rows[0]; idx = {h: i for i, h in enumerate(hdr)}
    want = ['gpu__time_duration.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
            'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
            'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
            'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
            'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size',
            'launch__block_size', 'sm__cycles_elapsed.max']
    want = [w for w in want if w in idx]
    with open(out, 'w') as f:
        w = csv.writer(f)
        w.writerow(['kernel'] + want)
        w.writerow(['(unit)'] + [rows[1][idx[c]] for c in want])
        for r in rows[2:]:
            name = r[idx['Kernel Name']]
            m = re.search(r'k_tc<cffm::tc::(\w+)<?([^>]*)', name)
            short = ('k_tc<' + m.group(1) + '<' + m.group(2) + '>>') if m else name.split('(')[0][:70]
            w.writerow([short] + [r[idx[c]] for c in want])

if __name__ == '__main__':
    kind, src, dst = sys.argv[1:4]
    (launch_summary if kind == 'launches' else full_summary)(src, dst)
