#!/usr/bin/env python
"""Mean per-role cycle counters written by MMG_SCAN_DBG_CLOCKS (see scan_tc_launch in api.cu)."""
import sys
import numpy as np
for path in sys.argv[1:]:
    try:
        a = np.loadtxt(path, comments='#')
    except Exception as e:
        print(path, 'unreadable', e)
        continue
    a = a[a[:, 1] > 0]
    mean = a.mean(axis=0) / 1e6
    names = ['cta', 'prod_total', 'prod_wait_empty', 'prod_wait_aempty', '-', 'mma_total', 'mma_wait_full', 'mma_wait_tempty', 'mma_wait_afull', 'epi_total', 'epi_wait_tfull', 'epi_xload', 'epi_fp64', 'epi_drain']
    print(path, ' '.join('%s=%.2f' % (n, v) for n, v in zip(names[1:], mean[1:]) if n != '-'), '(M cycles, %d CTAs)' % len(a))
