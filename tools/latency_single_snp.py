#!/usr/bin/env python
"""Host-side latency of the single-SNP scan calls of the stepwise / MLMM callers (linear_models.py:2720,2825):
`LinearMixedModel._emmax_f_test_([snp], H_sqrt_inv)` after one warm-up call -- wall time per call, the library's stage timers and
a cProfile of one call.  Usage: latency_single_snp.py [n] [m_for_kinship]"""
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
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
    m = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
    from mixmogam_b200 import _lib, kinship, linear_models as lm
    from oracle import reference_py3 as o
    ctx = _lib.get_context(0)
    snps = o.synth_genotypes(m, n, 'diploid_int', seed=99)
    K = kinship.calc_ibs_kinship(snps, 'diploid_int')
    y = np.random.default_rng(3).standard_normal(n) + 0.3 * snps[11]
    lmm = lm.LinearMixedModel(y)
    lmm.add_random_effect(K)
    lmm.add_factor(snps[7].astype(np.float64))
    eig_L = lmm._get_eigen_L_()
    res = lmm.get_estimates(eig_L=eig_L)
    for with_betas in (False, True):
        one = [snps[123]]
        lmm._emmax_f_test_(one, res['H_sqrt_inv'], emma_num=0, with_betas=with_betas)
        ts = []
        for _ in range(5):
            ctx.timer_reset()
            t0 = time.perf_counter()
            lmm._emmax_f_test_(one, res['H_sqrt_inv'], emma_num=0, with_betas=with_betas)
            ts.append(1e3 * (time.perf_counter() - t0))
        print('n=%d with_betas=%s: ms per call %s; stage timers of the last call (ms): %s' % (
            n, with_betas, [round(t, 2) for t in ts], {k: round(1e3 * v, 3) for k, v in ctx.timers().items() if v}))
        pr = cProfile.Profile()
        pr.enable()
        lmm._emmax_f_test_(one, res['H_sqrt_inv'], emma_num=0, with_betas=with_betas)
        pr.disable()
        buf = io.StringIO()
        pstats.Stats(pr, stream=buf).sort_stats('cumulative').print_stats(18)
        print(buf.getvalue()[:3500])


if __name__ == '__main__':
    main()
