#!/usr/bin/env python
"""BASELINE.json configs[0] (the reference's own CPU-runnable case: n = 198 accessions x ~214 000 binary SNPs, one phenotype):
wall time of kinship.calc_ibs_kinship + linear_models.emmax through the public API, stage timers and a cProfile of one call."""
import cProfile
import io
import os
import pstats
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 198
    m = int(sys.argv[2]) if len(sys.argv) > 2 else 214000
    from mixmogam_b200 import _lib, kinship, linear_models as lm
    from oracle import reference_py3 as o
    ctx = _lib.get_context(0)
    snps = o.synth_genotypes(m, n, 'binary', seed=1)
    y = np.random.default_rng(3).standard_normal(n) + 0.5 * snps[11]
    for rep in range(3):
        ctx.invalidate_snps()
        ctx.timer_reset()
        t0 = time.perf_counter()
        K = kinship.calc_ibs_kinship(snps, 'binary')
        t1 = time.perf_counter()
        r = lm.emmax(snps, y, K)
        t2 = time.perf_counter()
        print('n=%d m=%d rep %d: kinship %.1f ms, emmax %.1f ms; stage timers (ms): %s' % (
            n, m, rep, 1e3 * (t1 - t0), 1e3 * (t2 - t1), {k: round(1e3 * v, 2) for k, v in ctx.timers().items() if v}))
    pr = cProfile.Profile()
    ctx.invalidate_snps()
    pr.enable()
    K = kinship.calc_ibs_kinship(snps, 'binary')
    r = lm.emmax(snps, y, K)
    pr.disable()
    buf = io.StringIO()
    pstats.Stats(pr, stream=buf).sort_stats('cumulative').print_stats(30)
    print(buf.getvalue()[:6000])
    print('min p', float(np.min(r['ps'])), 'pseudo-heritability', r['pseudo_heritability'])


if __name__ == '__main__':
    main()
