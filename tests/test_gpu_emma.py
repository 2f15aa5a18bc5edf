"""GPU: exact EMMA for a batch of SNPs (mmg_emma_f64, csrc/emma.cuh) and the ML branch -- against the FP64 oracle, which
eigendecomposes S(K+I)S once per SNP exactly like the reference (linear_models.py:931-968, :771-927, :672-696), and against the
reference's own float32 run (tests/golden/ref_ml_emma_gxt_n400.npz)."""
import numpy as np
import pytest

from conftest import golden

pytestmark = pytest.mark.gpu


def _model(mod, y, K, cofs, **kw):
    lmm = mod.LinearMixedModel(list(y), **kw)
    lmm.add_random_effect(K)
    for c in cofs:
        lmm.add_factor(c)
    return lmm


@pytest.mark.parametrize('with_cof', [False, True])
def test_expedited_reml_batched_matches_oracle(ctx, with_cof):
    from mixmogam_b200 import linear_models as lm
    from oracle import reference_py3 as o
    e = golden('emmax_diploid_n400.npz')
    snps, y, K, cof = e['snps'], e['y'], e['K'], e['cofactor']
    cofs = [cof] if with_cof else []
    top = [snps[i] for i in (0, 1, 2, 3, 5, 8, 13, 21, 34, 55, 89, 144)]
    ro = _model(o, y, K, cofs, dtype='double').expedited_REML_t_test(top)
    mdl = _model(lm, y, K, cofs, ctx=ctx)
    r = mdl.expedited_REML_t_test(top)
    for k in ('f_stats', 'vgs', 'ves', 'var_perc', 'max_lls', 'rss'):
        np.testing.assert_allclose(r[k], ro[k], rtol=2e-6, atol=1e-9, err_msg=k)
    np.testing.assert_allclose(-np.log10(r['ps']), -np.log10(ro['ps']), rtol=1e-6, atol=1e-9)
    np.testing.assert_allclose(np.asarray(r['betas']), np.asarray(ro['betas']), rtol=1e-5, atol=1e-7)
    # the same SNPs taken from the resident genotype block (what emma_num > 0 does after a scan)
    ctx.invalidate_snps()
    ctx.ensure_snps(snps)
    rr = mdl.expedited_REML_t_test(None, _resident_rows=np.array([0, 1, 2, 3, 5, 8, 13, 21, 34, 55, 89, 144]))
    for k in ('ps', 'f_stats', 'vgs', 'max_lls'):
        np.testing.assert_array_equal(rr[k], r[k])
    # the reference's own float32 run, first 12 SNPs
    ref = golden('ref_ml_emma_gxt_n400.npz')
    tag = 'cof_' if with_cof else ''
    r12 = mdl.expedited_REML_t_test(list(snps[:12]))
    assert np.max(np.abs(np.log10(r12['ps']) - np.log10(ref['emma_' + tag + 'ps']))) < 2e-2
    np.testing.assert_allclose(r12['max_lls'], ref['emma_' + tag + 'max_lls'], atol=5e-2)
    ctx.invalidate_snps()


def test_get_estimates_with_xs_and_emma_num(ctx):
    """get_estimates(xs=snp) (one SNP, :771-927) and the emma_num refinement of emmax_f_test (:1365-1377) run through the same
    batch evaluator and agree with the oracle's eigendecomposition-per-SNP path."""
    from mixmogam_b200 import linear_models as lm
    from oracle import reference_py3 as o
    e = golden('emmax_diploid_n400.npz')
    snps, y, K = e['snps'][:1500], e['y'], e['K']
    mo = _model(o, y, K, [], dtype='double')
    eo = mo._get_eigen_L_()
    ro = mo.get_estimates(eo, xs=np.asarray(snps[7], dtype=np.float64).reshape(-1, 1), return_pvalue=True, return_f_stat=True)
    mdl = _model(lm, y, K, [], ctx=ctx)
    r = mdl.get_estimates(mdl._get_eigen_L_(), xs=np.asarray(snps[7], dtype=np.float64).reshape(-1, 1), return_pvalue=True,
                          return_f_stat=True)
    for k in ('delta', 'max_ll', 'vg', 've', 'f_stat', 'p_val', 'pseudo_heritability'):
        np.testing.assert_allclose(float(np.asarray(r[k]).reshape(-1)[0]), float(np.asarray(ro[k]).reshape(-1)[0]), rtol=2e-6, err_msg=k)
    np.testing.assert_allclose(np.asarray(r['beta']).reshape(-1), np.asarray(ro['beta']).reshape(-1), rtol=1e-5, atol=1e-8)
    ctx.invalidate_snps()
    full = lm.emmax(snps, y, K, emma_num=8, ctx=ctx)
    fo = o.emmax(list(snps), y, K, emma_num=8, dtype='double')
    a, b = -np.log10(full['ps']), -np.log10(fo['ps'])
    assert np.max(np.abs(a - b) / np.maximum(b, 1e-3)) < 1e-5
    ctx.invalidate_snps()


@pytest.mark.parametrize('with_cof', [False, True])
def test_get_ml_matches_oracle(ctx, with_cof):
    from mixmogam_b200 import linear_models as lm
    from oracle import reference_py3 as o
    e = golden('emmax_diploid_n400.npz')
    y, K, cof = e['y'], e['K'], e['cofactor']
    cofs = [cof] if with_cof else []
    ro = _model(o, y, K, cofs, dtype='double').get_ML()
    r = _model(lm, y, K, cofs, ctx=ctx).get_ML()
    for k in ('delta', 'max_ll', 'vg', 've', 'pseudo_heritability'):
        np.testing.assert_allclose(float(r[k]), float(ro[k]), rtol=2e-6, err_msg=k)
    np.testing.assert_allclose(np.asarray(r['beta']).reshape(-1), np.asarray(ro['beta']).reshape(-1), rtol=1e-6, atol=1e-9)
    np.testing.assert_allclose(np.asarray(r['mahalanobis_rss']).reshape(-1), np.asarray(ro['mahalanobis_rss']).reshape(-1), rtol=1e-6)
    ref = golden('ref_ml_emma_gxt_n400.npz')
    tag = 'cof_' if with_cof else ''
    assert abs(float(r['max_ll']) - float(ref['ml_' + tag + 'max_ll'])) < 5e-2         # the reference's float32 run
    # REML through the same eig_L-only evaluator == REML through eig_R (the path that mirrors the reference line by line)
    mdl = _model(lm, y, K, cofs, ctx=ctx)
    eig_L = mdl._get_eigen_L_()
    a = mdl.get_estimates(eig_L, method='REML')
    deltas = np.exp((np.arange(51, dtype=np.float64) / 50) * 20 - 10)
    b = ctx.emma(ctx.to_device(eig_L['vectors']), eig_L['values'], mdl.X, mdl.Y, deltas=deltas, method='REML')
    np.testing.assert_allclose(b['delta'][0], a['delta'], rtol=1e-8)
    np.testing.assert_allclose(b['max_ll'][0], a['max_ll'], rtol=1e-10)
    np.testing.assert_allclose(b['vg'][0], a['vg'], rtol=1e-8)
