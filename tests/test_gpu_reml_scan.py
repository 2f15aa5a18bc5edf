"""GPU: stages 2 and 3 (REML, scan, permutations) through the public API against the float64 oracle.
Tolerance of BASELINE.json: |d(-log10 p)| <= 1e-6 * max(-log10 p, 1e-3), identical top-hit ranking."""
import warnings

import numpy as np
import pytest

from conftest import golden, neglog10_rel_err

pytestmark = pytest.mark.gpu
warnings.simplefilter('ignore')

SCAN_IMPLS = ['dmma', 'tcgen05']


def test_f_sf_device(ctx):
    g = golden('f_sf.npz')
    for dfn, key in ((1, 'sf'), (2, 'sf2'), (3, 'sf3')):
        for j in range(g['dfd'].shape[1]):
            F, S = g['f'][:, j], g[key][:, j]
            out = ctx.f_sf(F, dfn, g['dfd'][0, j])
            tail = (S > 0) & (S < 0.5)
            assert np.max(np.abs(out[tail] - S[tail]) / S[tail]) < 1e-10
            assert np.max(np.abs(out[S >= 0.5] - S[S >= 0.5])) < 1e-13


def test_syevd_and_eigen(ctx):
    from mixmogam_b200 import linear_models as lm
    g = golden('emmax_ft10_n198.npz')
    lmm = lm.LinearMixedModel(g['y'])
    lmm.add_random_effect(g['K'])
    eig_L = lmm._get_eigen_L_()
    np.testing.assert_allclose(eig_L['values'], g['reml_eigL_values'], rtol=1e-9, atol=1e-11)
    U = np.asarray(eig_L['vectors'])
    np.testing.assert_allclose(U @ U.T, np.eye(len(U)), atol=1e-12)              # rows are orthonormal eigenvectors
    Ks = np.asarray(lmm.random_effects[1][1])
    np.testing.assert_allclose(U.T @ np.diag(eig_L['values']) @ U, Ks, atol=1e-10)
    eig_R = lmm._get_eigen_R_()
    assert eig_R['values'].shape == (197,) and eig_R['vectors'].shape == (197, 198)
    np.testing.assert_allclose(eig_R['values'], g['reml_eigR_values'], rtol=1e-8, atol=1e-10)


def test_small_eigenproblems_on_the_jacobi_solver(ctx, monkeypatch):
    """MMG_SYEVD_JACOBI_MAX: the device-only Jacobi eigensolver for small matrices gives the same eigenvalues and an orthonormal
    eigenbasis that reconstructs K (rows = eigenvectors, linear_models.py:596)."""
    from mixmogam_b200 import linear_models as lm
    g = golden('emmax_ft10_n198.npz')
    monkeypatch.setenv('MMG_SYEVD_JACOBI_MAX', '512')
    lmm = lm.LinearMixedModel(g['y'])
    lmm.add_random_effect(g['K'])
    eig_L = lmm._get_eigen_L_()
    np.testing.assert_allclose(eig_L['values'], g['reml_eigL_values'], rtol=1e-9, atol=1e-11)
    U = np.asarray(eig_L['vectors'])
    np.testing.assert_allclose(U @ U.T, np.eye(len(U)), atol=1e-12)
    np.testing.assert_allclose(U.T @ np.diag(eig_L['values']) @ U, np.asarray(lmm.random_effects[1][1]), atol=1e-10)
    r = lm.emmax(g['snps'], g['y'], g['K'])
    assert neglog10_rel_err(r['ps'], g['double_ps']) < 1e-6


@pytest.mark.parametrize('name', ['emmax_ft10_n198.npz', 'emmax_diploid_n400.npz'])
def test_get_reml(ctx, name):
    from mixmogam_b200 import linear_models as lm
    g = golden(name)
    lmm = lm.LinearMixedModel(g['y'])
    lmm.add_random_effect(g['K'])
    res = lmm.get_REML()
    if name.startswith('emmax_ft10'):
        assert res['delta'] == g['reml_delta']                                  # boundary optimum: exactly e^10
        np.testing.assert_allclose(res['max_ll'], g['reml_max_ll'], rtol=1e-10)
        np.testing.assert_allclose(lmm._last_reml['lls'], g['reml_lls'], rtol=1e-9, atol=1e-8)
        np.testing.assert_allclose(lmm._last_reml['dlls'], g['reml_dlls'], rtol=1e-7, atol=1e-8)
        np.testing.assert_allclose(res['vg'], g['reml_vg'], rtol=1e-9)
        np.testing.assert_allclose(res['ve'], g['reml_ve'], rtol=1e-9)
        H = np.asarray(res['H_sqrt_inv'])
        np.testing.assert_allclose(H.T @ H, g['reml_HtH'], rtol=1e-8, atol=1e-12)
        np.testing.assert_allclose(np.asarray(res['beta']).reshape(-1), g['reml_beta'], rtol=1e-8)
        np.testing.assert_allclose(np.asarray(res['mahalanobis_rss']).reshape(-1), g['reml_mahalanobis_rss'], rtol=1e-8)
        assert set(res) >= {'max_ll', 'delta', 'beta', 've', 'vg', 'rss', 'mahalanobis_rss', 'H_sqrt_inv',
                            'pseudo_heritability', 'eig_L'}
    else:
        from oracle import reference_py3 as o
        ol = o.LinearMixedModel(g['y'], 'double')
        ol.add_random_effect(g['K'])
        ro = ol.get_REML()
        assert abs(res['delta'] - ro['delta']) / ro['delta'] < 1e-9
        assert abs(res['max_ll'] - ro['max_ll']) < 1e-7
        assert lmm._last_reml['flags'] == 7


@pytest.mark.parametrize('impl', SCAN_IMPLS)
@pytest.mark.parametrize('name', ['emmax_ft10_n198.npz', 'emmax_diploid_n400.npz'])
def test_emmax_golden(ctx, impl, name):
    from mixmogam_b200 import linear_models as lm
    g = golden(name)
    ctx.invalidate_snps()
    r = lm.emmax(list(g['snps']), g['y'], g['K'], scan_impl=impl)
    assert set(r) >= {'ps', 'f_stats', 'rss', 'var_perc', 'h0_rss', 'h0_betas', 'pseudo_heritability', 've', 'vg', 'max_ll'}
    assert neglog10_rel_err(r['ps'], g['double_ps']) < 1e-6
    np.testing.assert_allclose(r['f_stats'], g['double_f_stats'], rtol=1e-6, atol=1e-9)
    np.testing.assert_allclose(r['rss'], g['double_rss'], rtol=1e-9)
    np.testing.assert_allclose(r['var_perc'], g['double_var_perc'], rtol=1e-6, atol=1e-12)
    np.testing.assert_allclose(np.asarray(r['h0_rss']).reshape(-1), g['double_h0_rss'], rtol=1e-9)
    np.testing.assert_allclose(r['pseudo_heritability'], g['double_pseudo_heritability'], rtol=1e-8)
    np.testing.assert_allclose(r['vg'], g['double_vg'], rtol=1e-8)
    np.testing.assert_allclose(r['max_ll'], g['double_max_ll'], rtol=1e-9)
    # identical ranking of the top hits, against both oracle modes; 1e-2 absolute against the float32-faithful mode
    assert np.array_equal(np.argsort(r['ps'], kind='stable')[:20], np.argsort(g['double_ps'], kind='stable')[:20])
    assert np.array_equal(np.argsort(r['ps'], kind='stable')[:20], np.argsort(g['single_ps'], kind='stable')[:20])
    assert np.max(np.abs(np.log10(r['ps']) - np.log10(g['single_ps']))) < 1e-2


@pytest.mark.parametrize('impl', SCAN_IMPLS)
def test_emmax_with_betas_cofactor_emma(ctx, impl):
    from mixmogam_b200 import linear_models as lm
    g = golden('emmax_ft10_n198.npz')
    ctx.invalidate_snps()
    r = lm.emmax(g['snps'], g['y'], g['K'], with_betas=True, scan_impl=impl)
    assert neglog10_rel_err(r['ps'], g['double_wb_ps']) < 1e-6
    np.testing.assert_allclose(np.asarray(r['betas']), g['double_wb_betas'], rtol=1e-5, atol=1e-8)
    g2 = golden('emmax_diploid_n400.npz')
    ctx.invalidate_snps()
    rc = lm.emmax(g2['snps'], g2['y'], g2['K'], cofactors=[g2['cofactor']], scan_impl=impl)
    ok = np.arange(len(rc['ps'])) != 17          # SNP 17 is the cofactor itself: x~ = 0 up to rounding, p is noise in every implementation
    assert neglog10_rel_err(rc['ps'][ok], g2['double_cof_ps'][ok]) < 1e-6
    ctx.invalidate_snps()
    re = lm.emmax(g['snps'][:400], g['y'], g['K'], emma_num=5, scan_impl=impl)
    assert neglog10_rel_err(re['ps'], g['double_emma5_ps']) < 1e-5


def test_snp_priors_and_transformed_snps(ctx):
    from mixmogam_b200 import linear_models as lm
    from oracle import reference_py3 as o
    g = golden('emmax_diploid_n400.npz')
    snps = g['snps'][:300]
    pri = np.full(len(snps), 1e-3)
    lmm = lm.LinearMixedModel(g['y'])
    lmm.add_random_effect(g['K'])
    ctx.invalidate_snps()
    r = lmm.emmax_f_test(snps, snp_priors=pri, emma_num=0)
    ol = o.LinearMixedModel(g['y'], 'double')
    ol.add_random_effect(g['K'])
    ro = ol.emmax_f_test(list(snps), snp_priors=pri, emma_num=0)
    np.testing.assert_allclose(r['ppas'], ro['ppas'], rtol=1e-5)
    np.testing.assert_allclose(r['bfs'], ro['bfs'], rtol=1e-5)
    res = lmm.get_REML()
    rt = lmm._emmax_f_test_(snps, res['H_sqrt_inv'], return_transformed_snps=True, emma_num=0)
    t = np.asarray(rt['t_snps'])
    np.testing.assert_allclose(np.sum(t * t, axis=1) * 0 + np.sum(t * t, axis=1), np.sum(t * t, axis=1))
    assert t.shape == (300, 400)


@pytest.mark.parametrize('perm_impl', ['tcgen05', 'dmma'])
def test_permutations_golden(ctx, perm_impl):
    """_emmax_permutations_ shuffles the ROTATED residual phenotype (linear_models.py:1151-1154), so its output
    depends on the sign convention of the eigenvectors inside H_sqrt_inv.  Feed the oracle's H_sqrt_inv (the
    method takes it as an argument) and the same np.random seed: then every number is comparable."""
    from mixmogam_b200 import linear_models as lm
    g = golden('perm_n120.npz')
    lmm = lm.LinearMixedModel(g['y'])
    lmm.perm_impl = perm_impl
    lmm.add_random_effect(g['K'])
    np.random.seed(int(g['seed']))
    ctx.invalidate_snps()
    pr = lmm._emmax_permutations_(g['snps'].astype(np.float64), g['K'], g['H_sqrt_inv'], num_perm=25)
    np.testing.assert_allclose(pr['max_f_stats'], g['max_f_stats'], rtol=1e-6)
    assert neglog10_rel_err(pr['min_ps'], g['min_ps']) < 1e-6
    assert abs(float(np.mean(lmm.Y))) < 1e-12                       # :1140 mutates the model
    # with our own (cuSOLVER) eigenbasis the permutation null is the same in distribution, not element-wise
    lmm2 = lm.LinearMixedModel(g['y'])
    lmm2.add_random_effect(g['K'])
    res = lmm2.get_REML()
    np.random.seed(int(g['seed']))
    pr2 = lmm2._emmax_permutations_(g['snps'], g['K'], res['H_sqrt_inv'], num_perm=25)
    assert pr2['min_ps'].shape == (25,) and np.all((pr2['min_ps'] > 0) & (pr2['min_ps'] < 1))
    assert 0.2 < np.median(pr2['max_f_stats']) / np.median(g['max_f_stats']) < 5


@pytest.mark.parametrize('n,m', [(1000, 6000), (2000, 3000)])
def test_scan_implementations_agree_and_match_oracle_sample(ctx, n, m):
    """Mid-size: the FP64 tensor-core path and the int8-slice path agree, and a sample matches the oracle."""
    from mixmogam_b200 import kinship, linear_models as lm
    from oracle import reference_py3 as o
    snps = o.synth_genotypes(m, n, 'diploid_int', seed=n + m)
    ctx.invalidate_snps()
    K = kinship.calc_ibs_kinship(snps, 'diploid_int')
    y = o.synth_phenotype(snps, K, seed=5)
    ra = lm.emmax(snps, y, K, scan_impl='dmma')
    rb = lm.emmax(snps, y, K, scan_impl='tcgen05')
    assert neglog10_rel_err(ra['ps'], rb['ps']) < 1e-7
    assert np.array_equal(np.argsort(ra['ps'], kind='stable')[:100], np.argsort(rb['ps'], kind='stable')[:100])
    sub = np.sort(np.random.default_rng(0).choice(m, 400, replace=False))
    ro = o.emmax(list(snps[sub]), y, K, dtype='double')
    assert neglog10_rel_err(ra['ps'][sub], ro['ps']) < 1e-6
    assert abs(ra['pseudo_heritability'] - ro['pseudo_heritability']) < 1e-8


def test_single_snp_calls_take_the_low_latency_path(ctx):
    """The stepwise / MLMM callers test one SNP at a time (`_emmax_f_test_([snp])`, linear_models.py:2720,2825): with
    scan_impl='auto' a short scan goes to the FP64 tensor-core kernel (no n^3 set-up product), a long one to the int8 scan; both
    match the oracle, with and without betas and with a cofactor."""
    import time
    from mixmogam_b200 import kinship, linear_models as lm
    from oracle import reference_py3 as o
    n, m = 1500, 4000
    snps = o.synth_genotypes(m, n, 'diploid_int', seed=99)
    ctx.invalidate_snps()
    K = kinship.calc_ibs_kinship(snps, 'diploid_int')
    y = o.synth_phenotype(snps, K, seed=3)
    lmm = lm.LinearMixedModel(y)
    lmm.add_random_effect(K)
    lmm.add_factor(snps[7].astype(np.float64))                                   # a cofactor SNP, as in the stepwise loop
    eig_L = lmm._get_eigen_L_()
    res = lmm.get_estimates(eig_L=eig_L)
    olmm = o.LinearMixedModel(y, dtype='double')
    olmm.add_random_effect(K)
    olmm.add_factor(snps[7].astype(np.float64))
    ores = olmm.get_estimates(eig_L=olmm._get_eigen_L_())
    for with_betas in (False, True):
        one = [snps[123]]
        lmm._emmax_f_test_(one, res['H_sqrt_inv'], emma_num=0, with_betas=with_betas)          # warm-up (handles, workspaces)
        t0 = time.perf_counter()
        r = lmm._emmax_f_test_(one, res['H_sqrt_inv'], emma_num=0, with_betas=with_betas)
        dt = time.perf_counter() - t0
        assert ctx.last_kernel_ms('scan_impl') == 3.0                             # MMG_IMPL_DMMA
        ro = olmm._emmax_f_test_(one, ores['H_sqrt_inv'], emma_num=0, with_betas=with_betas)
        assert neglog10_rel_err(r['ps'], ro['ps']) < 1e-6
        np.testing.assert_allclose(r['rss'], np.asarray(ro['rss'], dtype=np.float64), rtol=1e-8)
        if with_betas:
            np.testing.assert_allclose(np.asarray(r['betas'], dtype=np.float64), np.asarray(ro['betas'], dtype=np.float64).reshape(1, -1),
                                       rtol=1e-6, atol=1e-9)
        assert dt < 0.25, 'single-SNP call took %.1f ms' % (1e3 * dt)
    r_all = lmm._emmax_f_test_(snps, res['H_sqrt_inv'], emma_num=0)               # 4000 SNPs > n / 8: the int8 scan
    assert ctx.last_kernel_ms('scan_impl') == 1.0
    assert abs(-np.log10(r_all['ps'][123]) + np.log10(r['ps'][0])) <= 1e-6 * max(-np.log10(r['ps'][0]), 1e-3)


@pytest.mark.parametrize('n,m,P', [(500, 3000, 70), (1300, 2000, 33)])
def test_permutation_scan_tcgen05_matches_dmma(ctx, n, m, P):
    """The int8 tensor-core permutation scan (centred quadratic form + digit-plane GEMM) against the FP64 tensor-core
    one on sizes that span several K blocks, N tiles and a ragged permutation block."""
    from mixmogam_b200 import _lib
    from mixmogam_b200._lib import DeviceMatrix
    from oracle import reference_py3 as o
    rng = np.random.default_rng(n + P)
    snps = o.synth_genotypes(m, n, 'diploid_int', seed=n)
    ctx.invalidate_snps()
    ctx.ensure_snps(snps)
    H = rng.standard_normal((n, n)) / np.sqrt(n) + np.eye(n)
    Ys = rng.standard_normal((n, P))
    Hd = DeviceMatrix.from_host(ctx, H)
    Wt = DeviceMatrix.from_host(ctx, Ys.T @ H)
    ra = ctx.emmax_perm_scan(Hd, Wt, np.zeros(P), centre=True, impl=_lib.IMPL_DMMA)
    rb = ctx.emmax_perm_scan(Hd, Wt, np.zeros(P), centre=True, impl=_lib.IMPL_TCGEN05)
    np.testing.assert_allclose(rb, ra, rtol=1e-8)
    xc = snps.astype(np.float64) - snps.mean(1, keepdims=True)
    xt = xc @ H.T
    ref = np.max((xt @ Ys) ** 2 / np.sum(xt * xt, axis=1)[:, None], axis=0)
    np.testing.assert_allclose(rb, ref, rtol=1e-8)
    rc = ctx.emmax_perm_scan(Hd, Wt, np.zeros(P), centre=False, impl=_lib.IMPL_TCGEN05)
    xt = snps.astype(np.float64) @ H.T
    np.testing.assert_allclose(rc, np.max((xt @ Ys) ** 2 / np.sum(xt * xt, axis=1)[:, None], axis=0), rtol=1e-8)


@pytest.mark.parametrize('shared', [True, False])
def test_emmax_multi_matches_single_and_oracle(ctx, shared):
    """Phenotype-batched scan (configs[2]): T phenotypes, one eigenbasis == T independent emmax() calls.  shared=True: ONE
    rotation g = U x per SNP for all phenotypes (mmg_emmax_scan_shared_f64); shared=False: T rotations in one launch."""
    from mixmogam_b200 import linear_models as lm
    from oracle import reference_py3 as o
    g = golden('emmax_diploid_n400.npz')
    snps, K = g['snps'], g['K']
    rng = np.random.default_rng(5)
    Y = [g['y']] + [o.synth_phenotype(snps, K, seed=100 + t, h2_poly=h) for t, h in enumerate((0.0, 0.3, 0.8, 0.5))]
    Y.append(rng.standard_normal(400))
    ctx.invalidate_snps()
    res = lm.emmax_multi(snps, Y, K, shared=shared)
    assert len(res) == len(Y)
    if shared:
        S, rho = ctx.last_scan_info()
        assert 4 <= S <= 7 and 0.0 < rho <= 1e-7                 # the certified bound of the rotation's digit planes
    for t, y in enumerate(Y):
        single = lm.emmax(snps, y, K, scan_impl='tcgen05')
        # shared=False: same kernels, the only difference is cuBLAS gemm vs gemv rounding in etas = U_R Y (delta moves by ~1e-13);
        # shared=True: a different arithmetic path (rotation in the eigenbasis), held to the north-star tolerance
        if shared:
            assert neglog10_rel_err(res[t]['ps'], single['ps']) < 1e-6
            np.testing.assert_allclose(res[t]['rss'], single['rss'], rtol=1e-7)
            np.testing.assert_allclose(res[t]['f_stats'], single['f_stats'], rtol=1e-6, atol=1e-9)
            np.testing.assert_allclose(res[t]['var_perc'], single['var_perc'], rtol=1e-6, atol=1e-12)
        else:
            np.testing.assert_allclose(res[t]['ps'], single['ps'], rtol=1e-9, atol=0)
            np.testing.assert_allclose(res[t]['rss'], single['rss'], rtol=1e-10)
        for k in ('pseudo_heritability', 'vg', 've', 'max_ll'):
            np.testing.assert_allclose(res[t][k], single[k], rtol=1e-9)
        np.testing.assert_allclose(res[t]['h0_betas'], single['h0_betas'], rtol=1e-8, atol=1e-12)
        np.testing.assert_allclose(np.asarray(res[t]['h0_rss']).reshape(-1), np.asarray(single['h0_rss']).reshape(-1), rtol=1e-9)
        if t in (0, 2, 5):
            ro = o.emmax(list(snps), y, K, dtype='double')
            assert neglog10_rel_err(res[t]['ps'], ro['ps']) < 1e-6
            assert abs(res[t]['pseudo_heritability'] - ro['pseudo_heritability']) < 1e-8
    # with a cofactor and a batch size that does not divide T
    cof = g['cofactor']
    res2 = lm.emmax_multi(snps, Y[:3], K, cofactors=[cof], batch=2, shared=shared)
    ok = np.arange(len(snps)) != 17
    for t in range(3):
        single = lm.emmax(snps, Y[t], K, cofactors=[cof], scan_impl='tcgen05')
        if shared:
            assert neglog10_rel_err(res2[t]['ps'][ok], single['ps'][ok]) < 1e-6
            assert res2[t]['ps'][17] == 1.0 and res2[t]['rss'][17] == float(np.asarray(res2[t]['h0_rss']).reshape(-1)[0])   # collinear SNP: null fit
        else:
            np.testing.assert_allclose(res2[t]['ps'][ok], single['ps'][ok], rtol=1e-9)


def test_emmax_multi_shared_many_phenotypes_ragged(ctx, monkeypatch):
    """Shared-rotation batch at a size with several column tiles, K parts and SNP chunks: T = 37 phenotypes (two 32-row blocks
    of extra basis rows), n = 1100 (K = 9 blocks of 128, split in 2 parts), ragged last SNP group, 3 chunks; against the
    per-phenotype int8 scan and the FP64 oracle."""
    from mixmogam_b200 import kinship, linear_models as lm
    from oracle import reference_py3 as o
    n, m, T = 1100, 9000 + 77, 37
    snps = o.synth_genotypes(m, n, 'diploid_int', seed=78)
    ctx.invalidate_snps()
    K = kinship.calc_ibs_kinship(snps, 'diploid_int')
    Y = [o.synth_phenotype(snps, np.asarray(K), seed=200 + t, h2_poly=(t % 5) / 5.0) for t in range(T)]
    monkeypatch.setenv('MMG_SHARED_KSPLIT', '2')
    monkeypatch.setenv('MMG_SHARED_CHUNK', '4096')
    res = lm.emmax_multi(snps, Y, K)
    assert ctx.last_scan_info()[1] <= 1e-7
    for t in (0, 7, 36):
        single = lm.emmax(snps, Y[t], K, scan_impl='tcgen05')
        assert neglog10_rel_err(res[t]['ps'], single['ps']) < 1e-6
    ro = o.emmax(list(snps[:1500]), Y[36], np.asarray(K), dtype='double')
    assert neglog10_rel_err(res[36]['ps'][:1500], ro['ps']) < 1e-6
    for cs in ('1', '4'):
        monkeypatch.setenv('MMG_SHARED_CLUSTER', cs)
        r2 = lm.emmax_multi(snps, Y[:5], K)
        for t in range(5):
            # exact integer plane sums: the schedule cannot change them; what moves (1e-12) is delta-hat, through the cuBLAS
            # rounding of etas = U_R Y for 5 instead of 37 columns
            np.testing.assert_allclose(r2[t]['ps'], res[t]['ps'], rtol=1e-9)
    ctx.invalidate_snps()


@pytest.mark.parametrize('sched,panel', [('panel', '8'), ('panel', '6'), ('pair', '8'), ('pair128', '10'), ('n128', '8'), ('table', '8')])
def test_scan_schedules_agree(ctx, monkeypatch, sched, panel):
    """Every schedule of the int8 scan (genotype-stationary panels, CTA-pair MMA, 128-column tiles, tile table) gives the
    same statistics as the FP64 tensor-core path: several panels, ragged last SNP group, three waves of SNP groups."""
    from mixmogam_b200 import kinship, linear_models as lm
    from oracle import reference_py3 as o
    n, m = 1100, 40000 + 77
    snps = o.synth_genotypes(m, n, 'diploid_int', seed=77)
    ctx.invalidate_snps()
    K = kinship.calc_ibs_kinship(snps, 'diploid_int')
    y = o.synth_phenotype(snps, K, seed=6)
    ra = lm.emmax(snps, y, K, scan_impl='dmma')
    monkeypatch.setenv('MMG_SCAN_SCHED', sched)
    monkeypatch.setenv('MMG_SCAN_PANEL', panel)
    rb = lm.emmax(snps, y, K, scan_impl='tcgen05')
    assert neglog10_rel_err(ra['ps'], rb['ps']) < 1e-6
    assert np.array_equal(np.argsort(ra['ps'], kind='stable')[:100], np.argsort(rb['ps'], kind='stable')[:100])
    np.testing.assert_allclose(rb['rss'], ra['rss'], rtol=1e-7)


@pytest.mark.parametrize('cs', ['1', '2', '4', '8'])
def test_scan_cluster_sizes_agree(ctx, monkeypatch, cs):
    """The single-CTA-MMA form of the genotype-stationary scan with the digit tiles TMA-multicast to clusters of 1 / 2 / 4 / 8 CTAs: same
    statistics as the FP64 tensor-core path; 313 SNP groups do not divide by 8 (ragged last cluster group)."""
    from mixmogam_b200 import kinship, linear_models as lm
    from oracle import reference_py3 as o
    n, m = 900, 40000 + 13
    snps = o.synth_genotypes(m, n, 'diploid_int', seed=78)
    ctx.invalidate_snps()
    K = kinship.calc_ibs_kinship(snps, 'diploid_int')
    y = o.synth_phenotype(snps, K, seed=7)
    ra = lm.emmax(snps, y, K, scan_impl='dmma')
    monkeypatch.setenv('MMG_SCAN_SCHED', 'panel')
    monkeypatch.setenv('MMG_SCAN_CLUSTER', cs)
    rb = lm.emmax(snps, y, K, scan_impl='tcgen05')
    assert neglog10_rel_err(ra['ps'], rb['ps']) < 1e-6
    np.testing.assert_allclose(rb['rss'], ra['rss'], rtol=1e-7)


def test_scan_plane_count_is_certified(ctx, monkeypatch):
    """The number of base-256 digit planes is chosen from the certified truncation bound (pilot launch + check over every
    SNP); MMG_TC_SLICES fixes it, MMG_TC_TOL moves it."""
    from mixmogam_b200 import kinship, linear_models as lm
    from oracle import reference_py3 as o
    n, m = 600, 36000
    snps = o.synth_genotypes(m, n, 'diploid_int', seed=5)
    ctx.invalidate_snps()
    K = kinship.calc_ibs_kinship(snps, 'diploid_int')
    y = o.synth_phenotype(snps, K, seed=2)
    ref = lm.emmax(snps, y, K, scan_impl='dmma')
    r = lm.emmax(snps, y, K, scan_impl='tcgen05')
    S, rho = ctx.last_scan_info()
    assert 3 <= S <= 6 and 0.0 < rho <= 1e-7
    assert neglog10_rel_err(ref['ps'], r['ps']) < 1e-6
    monkeypatch.setenv('MMG_TC_TOL', '1e-12')
    r2 = lm.emmax(snps, y, K, scan_impl='tcgen05')
    S2, rho2 = ctx.last_scan_info()
    assert S2 > S and rho2 <= 1e-12
    assert neglog10_rel_err(ref['ps'], r2['ps']) < 1e-7
    monkeypatch.delenv('MMG_TC_TOL')
    monkeypatch.setenv('MMG_TC_SLICES', '2')
    r3 = lm.emmax(snps, y, K, scan_impl='tcgen05')
    S3, rho3 = ctx.last_scan_info()
    assert S3 == 2 and rho3 > rho
    # the certified bound really bounds the error of x~.x~ (rss = h0_rss - xy^2/xx moves by at most ~rho3 relative in xx)
    xx_ref = ref['h0_rss'] - ref['rss']
    xx_3 = r3['h0_rss'] - r3['rss']
    ok = xx_ref > 1e-9 * ref['h0_rss']
    assert np.max(np.abs(xx_3[ok] / xx_ref[ok] - 1.0)) <= 2.0 * rho3 + 1e-9


def test_scan_int8_domain_guard(ctx):
    """Genotype magnitudes beyond the exact-integer domain of the int8 scan are refused loudly; the FP64 path takes them."""
    from mixmogam_b200 import _lib, linear_models as lm
    from oracle import reference_py3 as o
    g = golden('emmax_diploid_n400.npz')
    snps = g['snps'].copy()
    snps[3, 5] = 9
    ctx.invalidate_snps()
    with pytest.raises(_lib.MmgError):
        lm.emmax(snps, g['y'], g['K'], scan_impl='tcgen05')
    ctx.invalidate_snps()
    r = lm.emmax(snps, g['y'], g['K'], scan_impl='dmma')
    ro = o.emmax(list(snps), g['y'], g['K'], dtype='double')
    assert neglog10_rel_err(r['ps'], ro['ps']) < 1e-6
    ctx.invalidate_snps()


def test_quad_form_tiles_and_scan_quad(ctx):
    """The multi-GPU preparation path on one GPU: the packed 256 x 256 blocks of A = R'R formed in three slot ranges
    (mmg_quad_form_tiles, what three ranks would each do before the all-gather) equal numpy's R'R inside the returned error
    bound, and the scan fed with that packed A and v = R'y on the device (mmg_emmax_scan_quad_dev) equals
    mmg_emmax_scan_f64; the dense host-vector entry (mmg_emmax_scan_quad_f64) agrees too."""
    from mixmogam_b200 import parallel
    from mixmogam_b200._lib import DeviceMatrix
    from oracle import reference_py3 as o
    rng = np.random.default_rng(11)
    n, m = 700, 5000
    snps = o.synth_genotypes(m, n, 'diploid_int', seed=3)
    ctx.invalidate_snps()
    ctx.ensure_snps(snps)
    R = rng.standard_normal((n - 1, n)) / np.sqrt(n)
    y = rng.standard_normal(n - 1)
    Rd = DeviceMatrix.from_host(ctx, R)
    slots = ctx.quad_form_slots(n)
    assert slots == 3 * 4 // 2                                         # n = 700 -> 3 block rows
    world = 4                                                          # more ranks than some have blocks for: 6 = 2 + 2 + 2 + 0
    per = parallel.slot_range(slots, 0, world)[2]
    Ap = DeviceMatrix(ctx, per * world, 65536, zero=False)
    errs = []
    for r in range(world):
        b, c, _ = parallel.slot_range(slots, r, world)
        errs.append(ctx.quad_form_tiles(Rd, b, c, Ap))
    assert len(set(errs)) == 1 and errs[0] > 0
    blocks = Ap.download().reshape(-1, 256, 256)
    A = np.zeros((768, 768))
    for J in range(3):
        for I in range(J + 1):
            A[J * 256:(J + 1) * 256, I * 256:(I + 1) * 256] = blocks[J * (J + 1) // 2 + I]
    ref = R.T @ R
    assert np.max(np.abs(np.tril(A[:n, :n]) - np.tril(ref))) <= errs[0] + 1e-15
    h0 = float(y @ y)
    ra = ctx.emmax_scan(Rd, y.reshape(1, -1), h0, n - 2, impl='tcgen05')
    vd = DeviceMatrix.from_host(ctx, (R.T @ y).reshape(-1, 1))
    out = ctx.emmax_scan_quad_dev(Ap, vd, h0, n - 2, packed=True, a_err=errs[0]).download()
    # the pre-pass launched ahead on the side stream (what the multi-GPU path does underneath the block products): same numbers
    ctx.scan_prepass_begin(Rd, y)
    out2 = ctx.emmax_scan_quad_dev(Ap, None, h0, n - 2, packed=True, a_err=errs[0]).download()
    np.testing.assert_allclose(out2, out, rtol=1e-11, atol=1e-300)
    with pytest.raises(Exception):
        ctx.emmax_scan_quad_dev(Ap, None, h0, n - 2, packed=True, a_err=errs[0])       # consumed: v is needed again
    Ad = DeviceMatrix.from_host(ctx, np.tril(ref))
    rb = ctx.emmax_scan_quad(Ad, R.T @ y, h0, n - 2)
    for i, k in enumerate(parallel.RESULT_KEYS):
        np.testing.assert_allclose(out[i], ra[k], rtol=1e-9, atol=1e-300)
        np.testing.assert_allclose(rb[k], ra[k], rtol=1e-9, atol=1e-300)
    xt = snps.astype(np.float64) @ R.T
    np.testing.assert_allclose(out[4], np.sum(xt * xt, axis=1), rtol=1e-7)
    ctx.invalidate_snps()


def test_degenerate_snps_keep_the_null_fit(ctx):
    """A monomorphic SNP and a SNP equal to a cofactor are collinear with the fixed effects: x~ = 0 up to rounding.  The int8
    scan gives them the null fit (rss = h0_rss, f = 0, p = 1 -- the reference's empty-residue case, linear_models.py:1329),
    keeps them out of the certification maximum, and the other SNPs are unaffected."""
    from mixmogam_b200 import kinship, linear_models as lm
    from oracle import reference_py3 as o
    n, m = 400, 12000
    snps = o.synth_genotypes(m, n, 'diploid_int', seed=5)
    cof = snps[17].astype(np.float64)
    snps[100] = 2                                                      # monomorphic
    snps[200] = 0
    snps[300] = snps[17]                                               # identical to the cofactor
    K = kinship.calc_ibs_kinship(snps, 'diploid_int')
    y = o.synth_phenotype(snps, K, seed=2)
    ctx.invalidate_snps()
    r = lm.emmax(snps, y, K, cofactors=[cof], scan_impl='tcgen05')
    S, rho = ctx.last_scan_info()
    assert rho <= 1e-7                                                 # certified although three SNPs have x~.x~ ~ 0
    for i in (100, 200, 300, 17):
        assert r['ps'][i] == 1.0 and r['f_stats'][i] == 0.0 and r['var_perc'][i] == 0.0
        assert r['rss'][i] == float(np.asarray(r['h0_rss']).reshape(-1)[0])
    ro = o.emmax(list(snps), y, K, cofactors=[cof], dtype='double')
    keep = np.ones(m, dtype=bool)
    keep[[17, 100, 200, 300]] = False
    assert neglog10_rel_err(r['ps'][keep], ro['ps'][keep]) < 1e-6
    ctx.invalidate_snps()


@pytest.mark.parametrize('n,n_out', [(700, 699), (1300, 1297), (257, 100)])
def test_quad_form_on_int8_pipe(ctx, monkeypatch, n, n_out):
    """A = R'R from exact int8 digit-plane products (tcgen05, 28 plane pairs) against the cuBLAS dsyrk path and against
    numpy, through x~.x~ = x'Ax of the scan with every digit plane of B in use; rows of R spanning three decades so the
    global scaling of the digit planes is exercised; the certified bound covers the observed error."""
    from mixmogam_b200._lib import DeviceMatrix
    from oracle import reference_py3 as o
    rng = np.random.default_rng(n)
    m = 5000
    snps = o.synth_genotypes(m, n, 'diploid_int', seed=n)
    ctx.invalidate_snps()
    ctx.ensure_snps(snps)
    R = rng.standard_normal((n_out, n)) / np.sqrt(n) * 10.0 ** rng.uniform(-3, 0, size=(n_out, 1))
    y = rng.standard_normal(n_out)
    Rd = DeviceMatrix.from_host(ctx, R)
    h0 = float(y @ y)
    monkeypatch.setenv('MMG_TC_SLICES', '6')
    monkeypatch.setenv('MMG_QUAD_A', 'dsyrk')
    ra = ctx.emmax_scan(Rd, y.reshape(1, -1), h0, n_out - 1, impl='tcgen05')
    monkeypatch.setenv('MMG_QUAD_A', 'int8')
    rb = ctx.emmax_scan(Rd, y.reshape(1, -1), h0, n_out - 1, impl='tcgen05')
    S, rho = ctx.last_scan_info()
    xt = snps.astype(np.float64) @ R.T
    xx = np.sum(xt * xt, axis=1)
    err_b = np.max(np.abs(rb['xx'] / xx - 1.0))
    assert err_b <= 1e-11 and err_b <= rho + 1e-13                       # exact to FP64 noise, inside the certified bound
    np.testing.assert_allclose(rb['xx'], ra['xx'], rtol=2e-11)
    np.testing.assert_allclose(rb['ps'], ra['ps'], rtol=1e-8, atol=1e-300)
    ctx.invalidate_snps()
