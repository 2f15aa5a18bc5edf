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


@pytest.mark.parametrize('with_cof', [False, True])
@pytest.mark.parametrize('impl', ['dmma', 'tcgen05'])
def test_emmax_w_two_env_matches_oracle_and_reference_run(ctx, with_cof, impl):
    """emmax_w_two_env -> _emmax_GxT_f_test_ (linear_models.py:1749-1787, :1422-1514): three tests per SNP from three passes of
    the fused scan (x, x o E, x o (1 + E)) against the FP64 oracle's two lstsq per SNP, and against the reference's own run."""
    from mixmogam_b200 import linear_models as lm
    from oracle import reference_py3 as o
    e = golden('emmax_diploid_n400.npz')
    ref = golden('ref_ml_emma_gxt_n400.npz')
    snps, K, cof = e['snps'][:600], e['K'], e['cofactor']
    E, ye = ref['E'], ref['ye']
    cofs = [cof] if with_cof else None
    ro = o.emmax_w_two_env(list(snps), list(ye), K, E, cofs, dtype='double')
    mdl = _model(lm, ye, K, cofs or [], ctx=ctx, scan_impl=impl)
    r = mdl.emmax_GxT_f_test(snps, E=E)
    tag = 'cof_' if with_cof else ''
    ok = np.ones(len(snps), dtype=bool)
    if with_cof:
        ok[17] = False          # SNP 17 is the cofactor itself: x~ = 0 up to rounding, every implementation's statistic is noise there
        h0 = float(np.asarray(r['g_res']['h0_rss']).reshape(-1)[0])
        assert r['g_res']['rss'][17] == h0 and r['gt_res']['rss'][17] == h0          # the product keeps the null fit (:1458)
    for part in ('g_res', 'gt_res', 'gt_g_res'):
        a, b = -np.log10(r[part]['ps'][ok]), -np.log10(ro[part]['ps'][ok])
        assert np.max(np.abs(a - b) / np.maximum(b, 1e-3)) < (1e-6 if impl == 'dmma' else 1e-5), part
        np.testing.assert_allclose(r[part]['f_stats'][ok], ro[part]['f_stats'][ok], rtol=1e-5, atol=1e-7, err_msg=part)
        np.testing.assert_allclose(r[part]['var_perc'][ok], ro[part]['var_perc'][ok], rtol=1e-5, atol=1e-9, err_msg=part)
        # the reference's float32 run (its three-column float32 lstsq is the noisy side: test_reference_pin holds the oracle to 5e-2 too)
        assert np.max(np.abs(np.log10(r[part]['ps']) - np.log10(ref['gxt_%s%s_ps' % (tag, part)]))[ok]) < 5e-2, part
    for part in ('g_res', 'gt_res'):
        np.testing.assert_allclose(r[part]['rss'][ok], np.asarray(ro[part]['rss']).reshape(-1)[ok], rtol=1e-8, err_msg=part)
        sel = np.flatnonzero(ok)        # (a SNP that kept the null fit carries h0_betas, a shorter list, like the reference's)
        np.testing.assert_allclose(np.asarray([r[part]['betas'][i] for i in sel]), np.asarray([ro[part]['betas'][i] for i in sel]),
                                   rtol=1e-5, atol=1e-7, err_msg=part)
    for k in ('pseudo_heritability', 've', 'vg', 'max_ll'):
        np.testing.assert_allclose(float(r[k]), float(ro[k]), rtol=2e-6, err_msg=k)
    if not with_cof and impl == 'dmma':
        top = lm.emmax_w_two_env(list(snps), list(ye), K, E, ctx=ctx)          # the module-level entry, list-of-rows input
        np.testing.assert_allclose(top['gt_g_res']['f_stats'], r['gt_g_res']['f_stats'], rtol=1e-9, atol=1e-12)


def test_emmax_w_two_env_degenerate_rows_keep_null_fit(ctx):
    """A SNP that never varies inside environment 1 gives a rank-deficient full model: the reference keeps h0_rss / h0_betas for
    it (:1463 `if rss_gt:`), a monomorphic SNP keeps them in both models (:1458)."""
    from mixmogam_b200 import linear_models as lm
    e = golden('emmax_diploid_n400.npz')
    ref = golden('ref_ml_emma_gxt_n400.npz')
    snps, K = np.array(e['snps'][:64]), e['K']
    E = ref['E']
    snps[3] = 1                                                               # monomorphic
    snps[5] = np.where(E[:, 0] > 0, 0, snps[5])                               # x o E == 0
    r = lm.emmax_w_two_env(snps, list(ref['ye']), K, E, ctx=ctx)
    h0 = float(np.asarray(r['g_res']['h0_rss']).reshape(-1)[0])
    assert r['g_res']['rss'][3] == h0 and r['gt_res']['rss'][3] == h0 and r['g_res']['ps'][3] == 1.0
    assert r['g_res']['rss'][5] < h0 and r['gt_res']['rss'][5] == h0
    assert r['gt_res']['betas'][5] == r['g_res']['h0_betas']


@pytest.mark.parametrize('with_betas', [False, True])
def test_real_valued_genotypes_fp64_rows(ctx, with_betas):
    """Imputed dosages (non-integral rows; linear_models.py:1317 casts whatever numeric row it gets): the FP64 tensor-core scan
    with the genotype operand staged as FP64, against the FP64 oracle; emma_num refinement on the same rows."""
    from mixmogam_b200 import linear_models as lm
    from oracle import reference_py3 as o
    e = golden('emmax_diploid_n400.npz')
    y, K = e['y'], e['K']
    rng = np.random.default_rng(11)
    xs = np.clip(e['snps'][:700].astype(np.float64) + rng.normal(0, 0.15, (700, 400)), 0.0, 2.0)
    ro = o.emmax([r for r in xs], y, K, with_betas=with_betas, emma_num=0, dtype='double')
    r = lm.emmax(xs, y, K, with_betas=with_betas, emma_num=0, ctx=ctx)
    a, b = -np.log10(r['ps']), -np.log10(ro['ps'])
    assert np.max(np.abs(a - b) / np.maximum(b, 1e-3)) < 1e-6
    np.testing.assert_allclose(r['rss'], np.asarray(ro['rss']).reshape(-1), rtol=1e-9)
    if with_betas:
        np.testing.assert_allclose(np.asarray(r['betas']), np.asarray(ro['betas']), rtol=1e-5, atol=1e-8)
    else:
        r5 = lm.emmax([row for row in xs], y, K, emma_num=5, ctx=ctx)         # list-of-rows input + refinement on dosages
        o5 = o.emmax([row for row in xs], y, K, emma_num=5, dtype='double')
        a, b = -np.log10(r5['ps']), -np.log10(o5['ps'])
        assert np.max(np.abs(a - b) / np.maximum(b, 1e-3)) < 1e-5
