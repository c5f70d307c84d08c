"""Summaries of ncu reports for profiles/: launch list (per-kernel mean duration + share) and the key raw
metrics of one `--set full` capture.  Usage: summarize.py launches <csv> | raw <ncu-rep>"""
import collections, csv, subprocess, sys

def launches(path):
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == 'ID')
    h = rows[hi]; ki, vi = h.index('Kernel Name'), h.index('Metric Value')
    agg = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) > vi:
            agg.setdefault(r[ki].split('(')[0][:70], []).append(float(r[vi].replace(',', '')))
    tot = sum(sum(v) for v in agg.values())
    print("%-72s %5s %12s %7s" % ("kernel", "n", "mean_ns", "share"))
    for k, v in agg.items():
        print("%-72s %5d %12.0f %6.1f%%" % (k, len(v), sum(v) / len(v), 100 * sum(v) / tot))

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_shared_mem', 'launch__waves_per_multiprocessor',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'lts__t_sector_hit_rate.pct']

def raw(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines())); h = rows[0]
    for k in KEYS:
        if k in h:
            i = h.index(k)
            print("%-62s %-14s %s" % (k, rows[1][i], [r[i] for r in rows[2:]]))

if __name__ == '__main__':
    {'launches': launches, 'raw': raw}[sys.argv[1]](sys.argv[2])
