"""GPU: BASELINE.json configs[1] at its FULL size (n = 10 000 individuals x 1 000 000 SNPs, diploid genotypes) checked
through size-independent properties -- the oracle cannot run there (the reference's own path needs ~30 min of CPU):

  kinship  * the streamed host Gram == the Gram of the resident block summed from two unequal SNP ranges (int32,
             bit for bit: additivity over SNPs, kinship.py:29-44)
           * trace(G) of the thermometer Gram == sum of all genotypes (every SNP counted exactly once)
           * unscaled K: symmetric, unit diagonal (kinship.py:51), entries in [0, 1]
           * a 3 x 1000 sample of the entries of the unscaled K == the reference's counts for those pairs (kinship.py:36-38,51)
             over all 1M SNPs, bit for bit
  REML     * max_ll at delta-hat == the restricted log-likelihood evaluated WITHOUT any eigendecomposition (Cholesky of
             K + delta I on the CPU: log|H|, log|X'H^-1 X|, y'Py) to 1e-9 relative, and delta-hat is a local maximum of that
             function -- pins cuSOLVER's eigenbases and the REML kernels at n = 10 000
  scan     * p-values of 400 sampled SNPs (top hits + random) == the GLS F-test computed from the same Cholesky factor
             (x'Px, x'Py, y'Py; linear_models.py:1272-1349 in its textbook form) to 1e-6 relative in -log10 p: an
             arithmetic path that shares nothing with the product (no eigenbasis, no rotation, no digit planes)
           * MMG_TEST_ORACLE_FULL=1 adds the line-faithful oracle itself (oracle.emmax, dtype='double': two n x n eigh on
             the CPU, ~5 min) on the same sample: p 1e-6, delta-hat 1e-9 (profiles/r02_parity_n10k.txt keeps one run)
           * the certified truncation bound of the int8 digit-plane scan holds over all 1M SNPs (<= 1e-7)
           * the int8 scan of the whole block == the FP64 tensor-core (DMMA) scan of a re-uploaded sample of its rows
             (top hits + random rows, shuffled) within 1e-6 relative in -log10 p, identical ranking of the top 100:
             row-order / batch independence and agreement of two independent arithmetic paths
           * allele flip x -> 2 - x leaves every p-value unchanged (the intercept is projected out,
             linear_models.py:1299-1303), within the same tolerance

Set MMG_TEST_FULL_M to run it on fewer SNPs (e.g. 131072) when iterating.
"""
import os

import numpy as np
import pytest

from conftest import neglog10_rel_err

pytestmark = pytest.mark.gpu

TOL = 1e-6          # relative in -log10 p (BASELINE.json north_star)


def test_full_size_properties(ctx):
    import torch
    import bench
    from mixmogam_b200 import kinship, linear_models as lm
    n = 10000
    m = int(os.environ.get('MMG_TEST_FULL_M', 1000000))
    dev = torch.device('cuda:0')
    snps = bench.gen_genotypes_pinned(0, m, n, dev)
    snps.flags.writeable = False                                         # residency contract: the device copy is reused across calls
    y = bench.gen_phenotype(n, dev)

    # ---- kinship ----
    ctx.invalidate_snps()
    assert ctx.kinship_gram_from(snps, 1) == (m, n)                      # streamed from host
    G = ctx.kinship_gram_download()
    cut = m // 2 + 77
    ctx.kinship_gram(1, snp_begin=0, snp_count=cut, reset=True)
    ctx.kinship_gram(1, snp_begin=cut, snp_count=m - cut, reset=False)
    assert np.array_equal(ctx.kinship_gram_download(), G)
    assert int(np.trace(G, dtype=np.int64)) == int(ctx.snps_row_sums().sum())
    del G
    Ku = np.asarray(kinship.calc_ibs_kinship(snps, 'diploid_int', scaled=False))
    assert np.array_equal(Ku, Ku.T) and np.all(np.diag(Ku) == 1.0) and Ku.min() >= 0.0 and Ku.max() <= 1.0
    from oracle import reference_py3 as o
    krng = np.random.default_rng(5)
    rows = krng.choice(n, size=3, replace=False)
    cols = np.sort(krng.choice(n, size=1000, replace=False))
    x_rows = np.ascontiguousarray(snps[:, rows].T).astype(np.int8)        # [3 x m]
    cnt = np.zeros((3, cols.size))
    for s0 in range(0, m, 100000):                                        # counts of equal / one-apart genotypes, SNP block by SNP block
        xb = np.ascontiguousarray(snps[s0:s0 + 100000][:, cols].T)        # [1000 x block] int8
        for a in range(3):
            d = np.abs(xb - x_rows[a][None, s0:s0 + 100000])
            cnt[a] += (d == 0).sum(axis=1, dtype=np.int64) + 0.5 * (d == 1).sum(axis=1, dtype=np.int64)      # kinship.py:36-38
    for a, r in enumerate(rows):
        ref = (cnt[a].astype(np.float32) / np.float32(m)).astype(np.float64)                                 # :51
        ref[cols == r] = 1.0                                                                                  # :35 (diagonal never filled) + I
        assert np.array_equal(Ku[r, cols], ref)
    del Ku, x_rows
    K = kinship.calc_ibs_kinship(snps, 'diploid_int')

    # ---- scan, whole block on the int8 tensor cores ----
    mdl = lm.LinearMixedModel(y, ctx=ctx, scan_impl='tcgen05')
    mdl.add_random_effect(K)
    eig_L = mdl._get_eigen_L_()
    eig_R = mdl._get_eigen_R_(X=mdl.X)
    full = mdl.emmax_f_test(snps, eig_L=eig_L, eig_R=eig_R, emma_num=0)
    S, rho = ctx.last_scan_info()
    assert 0.0 < rho <= 1e-7 and 3 <= S <= 6
    ps = full['ps']
    assert ps.shape == (m,) and np.all(np.isfinite(ps)) and ps.min() >= 0.0 and ps.max() <= 1.0   # causal SNPs underflow to 0
    assert 0.0 <= full['pseudo_heritability'] <= 1.0

    # ---- sample of rows: top hits + random rows, shuffled, through the FP64 tensor-core path ----
    rng = np.random.default_rng(1)
    fs = full['f_stats']                                                 # ranking by F: monotone in p, no underflow ties
    top = np.argsort(-fs, kind='stable')[:100]
    rest = np.setdiff1d(np.arange(m), top)                                 # the random part never repeats a top hit
    idx = np.concatenate([top, rng.choice(rest, size=3996, replace=False)])
    idx = idx[rng.permutation(idx.size)]
    sub = np.ascontiguousarray(snps[idx])
    ref = lm.LinearMixedModel(y, ctx=ctx, scan_impl='dmma')
    ref.add_random_effect(K)
    r_dmma = ref.emmax_f_test(sub, eig_L=eig_L, eig_R=eig_R, emma_num=0)
    assert neglog10_rel_err(ps[idx], r_dmma['ps']) < TOL
    order_full = idx[np.argsort(-fs[idx], kind='stable')[:100]]
    order_dmma = idx[np.argsort(-r_dmma['f_stats'], kind='stable')[:100]]
    assert np.array_equal(order_full, order_dmma) and np.array_equal(np.sort(order_full), np.sort(top))

    # ---- REML and scan against an eigendecomposition-free CPU evaluation (Cholesky of H = K + delta I) ----
    from scipy import linalg as sla
    Kh = np.array(np.asarray(mdl.random_effects[1][1]), dtype=np.float64)        # the scaled kinship the model holds (:580)
    delta = 1.0 / full['pseudo_heritability'] - 1.0
    X = np.ones((n, 1))
    yv = np.asarray(y, dtype=np.float64).reshape(-1, 1)
    pdim = n - 1

    def reml_parts(d, extra=None):
        H = Kh + d * np.eye(n)
        c = sla.cho_factor(H, lower=True, overwrite_a=True, check_finite=False)
        logdet = 2.0 * np.sum(np.log(np.diag(c[0])))
        B = np.hstack([X, yv] + ([extra] if extra is not None else []))
        Z = sla.cho_solve(c, B, check_finite=False)                              # H^-1 [X y xs]
        xhx = (X.T @ Z[:, :1]).item()
        py = Z[:, 1:2] - Z[:, :1] * ((X.T @ Z[:, 1:2]).item() / xhx)               # P y
        ypy = (yv.T @ py).item()
        ll = 0.5 * pdim * (np.log(pdim / (2.0 * np.pi)) - 1.0) - 0.5 * (pdim * np.log(ypy) + logdet + np.log(xhx) - np.log(float(n)))   # :618-623
        return ll, xhx, py, ypy, Z

    sidx = np.concatenate([top, rng.choice(rest, size=300, replace=False)])
    xs = np.ascontiguousarray(snps[sidx].T).astype(np.float64)                   # [n x 400]
    ll0, xhx, py, ypy, Z = reml_parts(delta, xs)
    assert abs(ll0 - full['max_ll']) <= 1e-9 * abs(ll0)
    for eps in (-2e-3, 2e-3):
        assert reml_parts(delta * (1.0 + eps))[0] <= ll0 + 1e-9 * abs(ll0)       # delta-hat is a local maximum of the REML likelihood
    Zx = Z[:, 2:]
    px = Zx - Z[:, :1] * ((X.T @ Zx) / xhx)                                      # P xs
    xpx = np.einsum('ij,ij->j', xs, px)
    xpy = (xs.T @ py).reshape(-1)
    r2 = xpy * xpy / (xpx * ypy)
    f_chol = (n - 2) * r2 / (1.0 - r2)                                           # :1346-1347 with rss = y'Py - (x'Py)^2 / x'Px
    from scipy import stats
    p_chol = stats.f.sf(f_chol, 1, n - 2)
    assert neglog10_rel_err(ps[sidx], p_chol) < TOL
    np.testing.assert_allclose(full['f_stats'][sidx], f_chol, rtol=1e-6)
    del Z, Zx, px

    if os.environ.get('MMG_TEST_ORACLE_FULL'):
        ro = o.emmax([np.asarray(r) for r in snps[sidx]], np.asarray(y), np.asarray(K), dtype='double')
        assert neglog10_rel_err(ps[sidx], ro['ps']) < TOL
        assert abs(ro['_delta'] - delta) <= 1e-9 * delta
        print('oracle(double) at n=%d on %d SNPs: max rel err in -log10 p %.3g, delta %.12g vs %.12g'
              % (n, sidx.size, neglog10_rel_err(ps[sidx], ro['ps']), delta, ro['_delta']))
    del Kh

    # ---- allele flip ----
    flipped = (2 - sub).astype(np.int8)
    r_flip = mdl.emmax_f_test(flipped, eig_L=eig_L, eig_R=eig_R, emma_num=0)
    assert ctx.last_scan_info()[1] <= 1e-7
    assert neglog10_rel_err(r_flip['ps'], ps[idx]) < TOL
    ctx.invalidate_snps()
