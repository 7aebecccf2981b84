"""Summarise ncu outputs: per-kernel launch times of the last step and key metrics of a full capture."""
import collections, csv, subprocess, sys

def launches(path):
    lines = [l for l in open(path) if not l.startswith('==')]
    rows = list(csv.DictReader(lines))
    idx = [i for i, x in enumerate(rows) if 'bow_score' in x['Kernel Name']]
    agg = collections.OrderedDict()
    for x in rows[idx[-1]:]:
        k = x['Kernel Name'][:50]; v = float(x['Metric Value'].replace(',', ''))
        u = x['Metric Unit']; v = v / 1e6 if u == 'ns' else v / 1e3 if u == 'us' else v
        agg.setdefault(k, []).append(v)
    tot = sum(sum(v) for v in agg.values())
    for k, v in agg.items():
        print("%-52s n=%2d total %8.3f ms (%4.1f%%) [%s]" % (k, len(v), sum(v), 100 * sum(v) / tot, ", ".join("%.2f" % t for t in v[:7])))
    print("total %.3f ms" % tot)

def metrics(rep):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.split('\n')))
    hdr = rows[0]
    want = ['gpu__time_duration.sum', 'smsp__inst_executed.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active',
            'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
            'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
            'smsp__thread_inst_executed_per_inst_executed.ratio', 'launch__registers_per_thread', 'launch__grid_size',
            'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
            'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active']
    for r in rows[2:]:
        if len(r) < len(hdr): continue
        print('--- kernel', r[hdr.index('Kernel Name')][:50], 'grid', r[hdr.index('Grid Size')], 'block', r[hdr.index('Block Size')])
        for w in want:
            if w in hdr: print('   %-70s %s' % (w, r[hdr.index(w)]))
        for i, hn in enumerate(hdr):
            if 'issue_stalled' in hn and 'per_issue_active' in hn:
                try:
                    v = float(r[i])
                except ValueError:
                    continue
                if v > 0.3: print('   stall %-64s %s' % (hn.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''), r[i]))

if __name__ == '__main__':
    if sys.argv[1].endswith('.csv'): launches(sys.argv[1])
    else: metrics(sys.argv[1])
