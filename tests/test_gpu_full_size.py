"""GPU: BASELINE.json configs[1] at its FULL size (n = 10 000 individuals x 1 000 000 SNPs, diploid genotypes) checked
through size-independent properties -- the oracle cannot run there (the reference's own path needs ~30 min of CPU):

  kinship  * the streamed host Gram == the Gram of the resident block summed from two unequal SNP ranges (int32,
             bit for bit: additivity over SNPs, kinship.py:29-44)
           * trace(G) of the thermometer Gram == sum of all genotypes (every SNP counted exactly once)
           * unscaled K: symmetric, unit diagonal (kinship.py:51), entries in [0, 1]
  scan     * the certified truncation bound of the int8 digit-plane scan holds over all 1M SNPs (<= 1e-7)
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
    del Ku
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
    idx = np.concatenate([top, rng.choice(m, size=3996, replace=False)])
    idx = idx[rng.permutation(idx.size)]
    sub = np.ascontiguousarray(snps[idx])
    ref = lm.LinearMixedModel(y, ctx=ctx, scan_impl='dmma')
    ref.add_random_effect(K)
    r_dmma = ref.emmax_f_test(sub, eig_L=eig_L, eig_R=eig_R, emma_num=0)
    assert neglog10_rel_err(ps[idx], r_dmma['ps']) < TOL
    order_full = idx[np.argsort(-fs[idx], kind='stable')[:100]]
    order_dmma = idx[np.argsort(-r_dmma['f_stats'], kind='stable')[:100]]
    assert np.array_equal(order_full, order_dmma) and np.array_equal(np.sort(order_full), np.sort(top))

    # ---- allele flip ----
    flipped = (2 - sub).astype(np.int8)
    r_flip = mdl.emmax_f_test(flipped, eig_L=eig_L, eig_R=eig_R, emma_num=0)
    assert ctx.last_scan_info()[1] <= 1e-7
    assert neglog10_rel_err(r_flip['ps'], ps[idx]) < TOL
    ctx.invalidate_snps()
