#!/usr/bin/env python
"""Summaries for profiles/: (1) launch list CSV (ncu --metrics gpu__time_duration.sum --csv) -> per-kernel share table,
(2) `ncu -i X.ncu-rep --page raw --csv` -> the metrics the roofline discussion uses."""
import csv
import sys
from collections import OrderedDict

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__m_xbar2l1tex_read_bytes.sum',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_subpipe_imma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active',
        'sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__cycles_elapsed.avg.per_second', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'launch__grid_size', 'launch__block_size', 'launch__cluster_size',
        'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic']


def launches(path):
    rows = [r for r in csv.reader(open(path)) if r and r[0].isdigit()]
    agg = OrderedDict()
    for r in rows:
        name, ns = r[4], float(r[-1])
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += ns
    tot = sum(v[1] for v in agg.values())
    print('# per-launch times are cold-cache and serialised: compare SHARES.  total = %.3f ms' % (tot / 1e6))
    print(' count      time_us   share  kernel')
    for name, (c, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print('%6d %12.1f  %5.1f%%  %s' % (c, ns / 1e3, 100 * ns / tot, name[:150]))


def raw(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        d = dict(zip(hdr, zip(units, vals)))
        print('--- %s' % d.get('Kernel Name', ('', '?'))[1][:160])
        for k in KEYS:
            if k in d:
                print('  %-80s %s %s' % (k, d[k][1], d[k][0]))


if __name__ == '__main__':
    {'launches': launches, 'raw': raw}[sys.argv[1]](sys.argv[2])
