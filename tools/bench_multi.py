#!/usr/bin/env python
"""
BASELINE.json configs[2] at the headline cohort size: T phenotypes (default 199) x m SNPs x n = 10 000 individuals scanned
against ONE kinship eigenbasis -- linear_models.emmax_multi with the shared rotation (mmg_emmax_scan_shared_f64) next to the
per-phenotype scans it replaces (T rotations in one launch, and single emmax() calls).  One JSON line: seconds per stage,
kernel milliseconds, SNP-tests/s, cost relative to one single-phenotype scan, parity of a few phenotypes.
Variants through the environment: MMG_SHARED_CLUSTER (1, 2, 4), MMG_SHARED_KSPLIT, MMG_SHARED_PLANES, MMG_SHARED_CHUNK.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mixmogam_b200 import _lib as _mmg_lib  # noqa: E402

_mmg_lib.load_library()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--indivs', dest='n', type=int, default=10000)
    ap.add_argument('--snps', dest='m', type=int, default=131072)
    ap.add_argument('--phenotypes', dest='T', type=int, default=199)
    ap.add_argument('--single', type=int, default=2)
    ap.add_argument('--unshared', type=int, default=4, help='phenotypes to run through the per-phenotype-rotation launch for comparison')
    ap.add_argument('--envs', default='', help='semicolon list of comma separated ENV=value settings to time the shared scan under')
    ap.add_argument('--variants', default='', help='comma list of cluster:ksplit pairs to time besides the default, e.g. 1:1,4:1,2:2')
    a = ap.parse_args()
    import torch
    import bench
    from mixmogam_b200 import _lib, kinship, linear_models as lm
    ctx = _lib.get_context(0)
    dev = torch.device('cuda:0')
    snps = bench.gen_genotypes_pinned(0, a.m, a.n, dev)
    snps.flags.writeable = False
    rng = np.random.Generator(np.random.PCG64(20240601 + 2))
    K = kinship.calc_ibs_kinship(snps, 'diploid_int')
    Y = []
    for t in range(a.T):
        idx = rng.choice(a.m, 300, replace=False)
        gval = rng.normal(0, 1.0, 300) @ (snps[np.sort(idx)].astype(np.float64))
        gval = (gval - gval.mean()) / gval.std()
        h2 = rng.uniform(0.0, 0.9)
        y = np.sqrt(h2) * gval + np.sqrt(1 - h2) * rng.standard_normal(a.n)
        Y.append((y - y.mean()) / y.std())
    lm.emmax_multi(snps, Y[:4], K)                             # warm-up: handles, kernels, workspaces (eigh of K twice)
    ctx.timer_reset()
    t0 = time.perf_counter()
    res = lm.emmax_multi(snps, Y, K)
    t_multi = time.perf_counter() - t0
    timers = ctx.timers()
    planes, rho = ctx.last_scan_info()
    line = {'config': 'configs[2] multi-phenotype batch at n = %d' % a.n, 'n': a.n, 'm': a.m, 'T': a.T,
            'emmax_multi_s': t_multi, 'stage_seconds': timers, 'planes': planes, 'certified_rel_bound_xx': rho,
            'scan_kernels_ms': ctx.last_kernel_ms('scan'), 'shared_scan_info': getattr(ctx, 'last_shared_info', None),
            'snp_tests_per_s_scan_stage': a.m * a.T / max(timers['scan'], 1e-9),
            'snp_tests_per_s_whole_call_excl_eigh': a.m * a.T / max(t_multi - timers['syevd'], 1e-9)}
    # single-phenotype scans (what the reference does T times, linear_models.py:1790)
    errs = []
    ts = []
    for t in range(min(a.single, a.T)):
        ctx.timer_reset()
        t0 = time.perf_counter()
        r1 = lm.emmax(snps, Y[t], K)
        ts.append((time.perf_counter() - t0, ctx.timers()['scan'], ctx.timers()['syevd']))
        lp, lq = -np.log10(np.maximum(res[t]['ps'], 1e-300)), -np.log10(np.maximum(r1['ps'], 1e-300))
        errs.append(float(np.max(np.abs(lp - lq) / np.maximum(lq, 1e-3))))
    line['single_emmax_s'] = [x[0] for x in ts]
    line['single_scan_stage_s'] = [x[1] for x in ts]
    line['max_rel_err_neglog10p_vs_single'] = errs
    if ts:
        line['scan_cost_vs_one_single_scan'] = timers['scan'] / max(min(x[1] for x in ts), 1e-9)
    if a.unshared:
        ctx.timer_reset()
        t0 = time.perf_counter()
        lm.emmax_multi(snps, Y[:a.unshared], K, shared=False)
        line['unshared_T%d_s' % a.unshared] = time.perf_counter() - t0
        line['unshared_scan_stage_s_per_phenotype'] = ctx.timers()['scan'] / a.unshared
    for v in [x for x in a.variants.split(',') if x]:
        cs, ks = v.split(':')
        os.environ['MMG_SHARED_CLUSTER'], os.environ['MMG_SHARED_KSPLIT'] = cs, ks
        ctx.timer_reset()
        lm.emmax_multi(snps, Y, K)
        line['variant_cluster%s_ksplit%s_scan_stage_s' % (cs, ks)] = ctx.timers()['scan']
        line['variant_cluster%s_ksplit%s_info' % (cs, ks)] = getattr(ctx, 'last_shared_info', None)
    for ev in [x for x in a.envs.split(';') if x]:
        sets = dict(kv.split('=') for kv in ev.split(','))
        old = {k: os.environ.get(k) for k in sets}
        os.environ.update(sets)
        ctx.timer_reset()
        lm.emmax_multi(snps, Y, K)
        line['env[%s]' % ev] = {'scan_stage_s': ctx.timers()['scan'], 'info': getattr(ctx, 'last_shared_info', None)}
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    print(json.dumps(line))


if __name__ == '__main__':
    main()
