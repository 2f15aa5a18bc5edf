#!/usr/bin/env python
"""
Timings of the BASELINE.json configurations that are NOT the bench.py headline (configs[1]):

  --config 2 : multi-phenotype batch -- T phenotypes (default 199) on n=198 accessions x 214k binary SNPs,
               one kinship eigenbasis, linear_models.emmax_multi (one scan launch) vs T x linear_models.emmax
  --config 4 : hdf5_data.run_emmax_perm -- n=5000 x 1M diploid SNPs in 5 chromosomes, IBD kinship with the MAF
               filter, EMMAX scan of every chromosome, 1000 permuted phenotypes on all chromosomes but the last
  --config 3 : large cohort n x m (default 50000 x 500000 takes 25 GB of pinned host memory; use --indivs/--snps)

Prints one JSON line per configuration (wall-clock seconds per stage, CUDA-event kernel times where the library
records them).  Synthetic data, seeds as SURVEY.md 8d.  Not the driver's bench: bench.py is.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
# libmixmogam_b200.so is loaded BEFORE torch: the process then resolves libcusolver.so.11 to the CUDA 12.9 toolkit copy the library
# was linked against (its Xsyevd takes n <= ~32768; the older copy bundled with the torch wheel refuses smaller matrices still)
from mixmogam_b200 import _lib as _mmg_lib  # noqa: E402

_mmg_lib.load_library()


def gen_torch(m, n, seed, binary, device, maf_lo=0.11):
    import torch
    from mixmogam_b200 import _lib
    host = _lib.pinned_empty((m, n), np.int8)
    th = torch.from_numpy(host)
    step = 32768
    for c, r0 in enumerate(range(0, m, step)):
        rows = min(step, m - r0)
        g = torch.Generator(device=device)
        g.manual_seed(seed * 1000003 + c)
        f = maf_lo + (0.5 - maf_lo) * torch.rand((rows, 1), generator=g, device=device)
        a = (torch.rand((rows, n), generator=g, device=device) < f).to(torch.int8)
        if not binary:
            a = a + (torch.rand((rows, n), generator=g, device=device) < f).to(torch.int8)
        th[r0:r0 + rows].copy_(a)
    torch.cuda.synchronize()
    return host


def config2(args):
    import torch
    from mixmogam_b200 import _lib, kinship, linear_models as lm
    ctx = _lib.get_context(0)
    n, m, T = args.n or 198, args.m or 214000, args.T
    dev = torch.device('cuda:0')
    snps = gen_torch(m, n, 20240601 + 2, True, dev, maf_lo=0.05)
    keep = snps.min(1) != snps.max(1)
    snps = np.ascontiguousarray(snps[keep])
    rng = np.random.Generator(np.random.PCG64(20240601 + 2))
    t0 = time.perf_counter()
    K = kinship.calc_ibs_kinship(snps, 'binary')
    t_kin = time.perf_counter() - t0
    L = np.linalg.cholesky(np.asarray(K) + 1e-6 * np.eye(n))
    Y = []
    for t in range(T):
        causal = rng.choice(len(snps), 5, replace=False)
        h2 = rng.uniform(0.0, 0.8)
        y = rng.normal(0, 0.5, 5) @ snps[causal] + np.sqrt(h2) * (L @ rng.standard_normal(n)) + np.sqrt(1 - h2) * rng.standard_normal(n)
        Y.append((y - y.mean()) / y.std())
    lm.emmax_multi(snps[:4096], Y[:2], K)                      # warm-up (cuSOLVER / cuBLAS handles, kernels)
    ctx.timer_reset()
    t0 = time.perf_counter()
    res = lm.emmax_multi(snps, Y, K)
    t_multi = time.perf_counter() - t0
    scan_ms = ctx.last_kernel_ms('scan')
    timers = ctx.timers()
    Tsub = min(T, args.single)
    t0 = time.perf_counter()
    singles = [lm.emmax(snps, Y[t], K) for t in range(Tsub)]
    t_single = (time.perf_counter() - t0) / Tsub
    err = max(float(np.max(np.abs(np.log10(res[t]['ps']) - np.log10(singles[t]['ps'])))) for t in range(Tsub))
    print(json.dumps({'config': 'configs[2] multi-phenotype batch', 'n': n, 'm': int(len(snps)), 'T': T,
                      'kinship_s': t_kin, 'emmax_multi_s': t_multi, 'scan_kernel_ms': scan_ms,
                      'snp_tests_per_s': len(snps) * T / t_multi, 'emmax_single_s_per_phenotype': t_single,
                      'speedup_vs_T_single_calls': t_single * T / t_multi, 'max_abs_dlog10p_vs_single': err,
                      'stage_seconds': timers}))


def config4(args):
    import torch
    from mixmogam_b200 import _lib, hdf5_data
    ctx = _lib.get_context(0)
    n, m, P, C = args.n or 5000, args.m or 1000000, args.P, 5
    dev = torch.device('cuda:0')
    rng = np.random.Generator(np.random.PCG64(20240601 + 4))
    per = m // C
    gg = {}
    for c in range(C):
        x = gen_torch(per, n, 20240601 + 4 + 17 * c, False, dev, maf_lo=0.05)
        gg['chrom_%d' % (c + 1)] = {'raw_snps': x, 'freqs': x.mean(1, dtype=np.float64) / 2.0, 'positions': np.arange(per) * 100 + 1}
    x0 = gg['chrom_1']['raw_snps']
    y = rng.normal(0, 0.5, 10) @ x0[:10] + rng.standard_normal(n)
    f = {'genot_data': gg, 'indiv_data': {'indiv_ids': np.arange(n), 'phenotypes': (y - y.mean()) / y.std()}, 'num_snps': np.array(m)}
    small = {'genot_data': {k: {kk: vv[:4096] for kk, vv in v.items()} for k, v in list(gg.items())[:2]},
             'indiv_data': f['indiv_data'], 'num_snps': np.array(8192)}
    np.random.seed(1)
    hdf5_data.run_emmax_perm(small, {}, num_perm=8)            # warm-up
    ctx.timer_reset()
    out = {}
    np.random.seed(20240601 + 4)
    t0 = time.perf_counter()
    hdf5_data.run_emmax_perm(f, out, min_maf=0.1, num_perm=P)
    t_all = time.perf_counter() - t0
    timers = ctx.timers()
    print(json.dumps({'config': 'configs[4] hdf5_data.run_emmax_perm', 'n': n, 'm': m, 'chromosomes': C, 'num_perm': P,
                      'total_s': t_all, 'perm_kernel_ms_last_call': ctx.last_kernel_ms('perm'), 'ibd_kernel_ms_last_chrom': ctx.last_kernel_ms('ibd'),
                      'five_perc_perm_min_ps': float(out['five_perc_perm_min_ps']),
                      'perm_tests_per_s': (m - per) * P / max(timers.get('scan', 0.0), 1e-9), 'stage_seconds': timers}))


def standin_eigen(ctx, n):
    """Orthonormal stand-in for the two eigenbases where cuSOLVER's syevd refuses the size (n > 32768): ONE Householder reflection
    H = I - 2 v v' with H e_1 = 1/sqrt(n) 1, built on the device.  eig_L = (ascending positive values, rows of H); eig_R = rows 1..n-1
    of the same matrix (orthogonal to the intercept, as linear_models.py:600-615 requires) with the values shifted by one place.
    The numbers that come out of the scan mean nothing; the kernels, their operands and their memory are the real ones."""
    import torch
    from mixmogam_b200 import parallel
    from mixmogam_b200._lib import DeviceMatrix, LazyHostArray
    from mixmogam_b200.linear_models import EigenDict
    dev = 'cuda:%d' % ctx.device
    v = torch.full((n,), -1.0 / np.sqrt(n), dtype=torch.float64, device=dev)
    v[0] += 1.0
    v = v / v.norm()
    U = DeviceMatrix(ctx, n, n, zero=False)
    ctx.sync()
    t = parallel.mat_as_tensor(ctx, U)
    step = 2048
    for r0 in range(0, n, step):
        r1 = min(n, r0 + step)
        blk = -2.0 * torch.outer(v[r0:r1], v)
        blk[torch.arange(r1 - r0, device=dev), torch.arange(r0, r1, device=dev)] += 1.0
        t[r0:r1, :n].copy_(blk)
    torch.cuda.synchronize()
    del blk, t
    torch.cuda.empty_cache()
    w = np.linspace(0.2, 3.0, n)
    return (EigenDict(values=w, vectors=LazyHostArray(U)),
            EigenDict(values=w[1:].copy(), vectors=LazyHostArray(U, rows=(1, n)), _q=1))


def config3(args):
    import torch
    from mixmogam_b200 import _lib, kinship, linear_models as lm
    ctx = _lib.get_context(0)
    n, m = args.n or 50000, args.m or 500000
    dev = torch.device('cuda:0')
    snps = gen_torch(m, n, 20240601 + 3, False, dev)
    torch.cuda.empty_cache()                                   # the generator's temporaries go back to the driver: the library needs the room
    rng = np.random.Generator(np.random.PCG64(20240601 + 3))
    y = rng.normal(0, 0.5, 10) @ snps[:10] + rng.standard_normal(n)
    ctx.timer_reset()
    t0 = time.perf_counter()
    K = kinship.calc_ibs_kinship_device(snps, 'diploid_int')
    ctx.sync()
    t_kin = time.perf_counter() - t0
    mdl = lm.LinearMixedModel(y, ctx=ctx)
    mdl.add_random_effect(K)
    K.free()
    t0 = time.perf_counter()
    if args.standin_eigen:
        eig_L, eig_R = standin_eigen(ctx, n)
    else:
        eig_L = mdl._get_eigen_L_()
        eig_R = mdl._get_eigen_R_(X=mdl.X)
    t_eig = time.perf_counter() - t0
    t0 = time.perf_counter()
    r = mdl.emmax_f_test(snps, eig_L=eig_L, eig_R=eig_R, emma_num=0)
    t_scan = time.perf_counter() - t0
    planes, rho = ctx.last_scan_info()
    info = ctx.device_info()
    print(json.dumps({'config': 'configs[3] large cohort', 'n': n, 'm': m, 'kinship_s': t_kin, 'gram_kernel_ms': ctx.last_kernel_ms('gram'),
                      'eigenbases': 'orthonormal stand-in (Householder reflection built on the device): cuSOLVER syevd refuses n > 32768'
                                    if args.standin_eigen else 'cuSOLVER Xsyevd', 'eigh_s': t_eig,
                      'reml_scan_s': t_scan, 'scan_kernel_ms': ctx.last_kernel_ms('scan'), 'planes': planes, 'certified_rel_bound_xx': rho,
                      'snp_tests_per_s_excl_eigh': m / (t_kin + t_scan), 'min_p': float(np.min(r['ps'])), 'finite': bool(np.all(np.isfinite(r['ps']))),
                      'pseudo_heritability': float(r['pseudo_heritability']), 'device_free_bytes_after': info.get('free_bytes'),
                      'stage_seconds': ctx.timers()}))


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--config', type=int, required=True, choices=[2, 3, 4])
    ap.add_argument('--indivs', dest='n', type=int, default=0)
    ap.add_argument('--snps', dest='m', type=int, default=0)
    ap.add_argument('--phenotypes', dest='T', type=int, default=199)
    ap.add_argument('--perms', dest='P', type=int, default=1000)
    ap.add_argument('--standin-eigen', action='store_true', help='config 3: orthonormal stand-in eigenbases instead of cuSOLVER (n > 32768)')
    ap.add_argument('--single', type=int, default=5, help='config 2: how many single-phenotype emmax() calls to time for the comparison')
    a = ap.parse_args()
    {2: config2, 3: config3, 4: config4}[a.config](a)
