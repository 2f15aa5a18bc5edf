"""GPU: the CUDA path against outputs of the REFERENCE'S OWN CODE (tests/golden/ref_*.npz, produced by
tests/golden/make_reference_golden.py from the unmodified reference sources).

The reference computes in float32 (linear_models.py:558,589,600,773,1283); the device computes the same algebra
in float64.  Tolerances are therefore the reference's own float32 noise, as SURVEY.md 8c states them:
kinship bit-exact; |d(-log10 p)| <= 1e-2 absolute; identical top-20 ranking; variance components to 1e-4.
(The 1e-6 criterion is held against the float64 oracle in test_gpu_reml_scan.py; the oracle itself reproduces
these reference runs bit for bit, tests/test_reference_pin.py.)
"""
import warnings

import numpy as np
import pytest

from conftest import golden, neglog10_rel_err

pytestmark = pytest.mark.gpu
warnings.simplefilter('ignore')


def assert_ps_close(r, ref, prefix='', top=20, skip=None):
    a = -np.log10(np.asarray(r['ps'], dtype=np.float64))
    b = -np.log10(np.asarray(ref[prefix + 'ps'], dtype=np.float64))
    if skip is not None:
        a, b = a[skip], b[skip]
    assert np.max(np.abs(a - b)) <= 1e-2
    assert np.array_equal(np.argsort(-a, kind='stable')[:top], np.argsort(-b, kind='stable')[:top])


def ref_delta(ref, prefix=''):
    return 1.0 / float(ref[prefix + 'pseudo_heritability']) - 1.0


def assert_reml_consistent(lmm, r, ref, prefix=''):
    """delta is only comparable where the likelihood is curved and the reference's float32 secant converged: the
    reference picks a grid point when its float32 refinement fails (always, for a flat likelihood; and under NEP-50
    numpy whenever float32 cannot meet tol=1e-6 -- oracle promotion='numpy2').  What must hold in every case: on the
    same restricted likelihood, our optimum is at least as good as the reference's."""
    d_ref, d_our = ref_delta(ref, prefix), 1.0 / float(r['pseudo_heritability']) - 1.0
    eig_R = lmm._last_reml['eig_R']
    ev = np.asarray(eig_R['values'], dtype=np.float64)
    sq = np.asarray(lmm._etas(eig_R, lmm.Y)) ** 2
    assert lmm._rell_(d_our, ev, sq) >= lmm._rell_(d_ref, ev, sq) - 1e-6
    if abs(d_our / d_ref - 1.0) < 1e-3:
        for k in ('ve', 'vg', 'max_ll'):
            np.testing.assert_allclose(float(r[k]), float(ref[prefix + k]), rtol=2e-3)
    return d_ref


def scan_at_reference_delta(lmm, snps, ref, prefix='', skip=None, **kw):
    """The scan with the REFERENCE's delta plugged in (H_sqrt_inv is an argument of _emmax_f_test_, :1272): every
    output is then comparable at float32 accuracy."""
    ctx = lmm.ctx
    eig_L = lmm._get_eigen_L_()
    H = ctx.to_device(eig_L['vectors']).copy()
    ctx.scale_rows(H, 1.0 / np.sqrt(np.asarray(eig_L['values'], dtype=np.float64) + ref_delta(ref, prefix)))
    r = lmm._emmax_f_test_(snps, H, emma_num=0, **kw)
    sel = slice(None) if skip is None else skip
    a = -np.log10(np.asarray(r['ps'], dtype=np.float64))[sel]
    b = -np.log10(ref[prefix + 'ps'])[sel]
    assert np.max(np.abs(a - b)) <= 5e-3
    assert np.array_equal(np.argsort(-a, kind='stable')[:20], np.argsort(-b, kind='stable')[:20])
    np.testing.assert_allclose(np.asarray(r['rss'], dtype=np.float64)[sel], ref[prefix + 'rss'][sel], rtol=1e-4)
    np.testing.assert_allclose(np.asarray(r['var_perc'], dtype=np.float64)[sel], ref[prefix + 'var_perc'][sel], atol=2e-5)
    np.testing.assert_allclose(np.asarray(r['f_stats'], dtype=np.float64)[sel], ref[prefix + 'f_stats'][sel], rtol=2e-2, atol=5e-3)
    np.testing.assert_allclose(np.asarray(r['h0_rss'], dtype=np.float64).reshape(-1), ref[prefix + 'h0_rss'], rtol=1e-4)
    np.testing.assert_allclose(np.asarray(r['h0_betas'], dtype=np.float64), ref[prefix + 'h0_betas'], rtol=1e-3, atol=1e-5)
    return r


def test_kinship_vs_reference_run(ctx):
    from mixmogam_b200 import kinship
    ref = golden('ref_kinship_n37.npz')
    xb = golden('ibs_binary_n37.npz')['snps']
    xd = golden('ibs_diploid_n37.npz')['snps']
    assert np.array_equal(np.asarray(kinship.calc_ibs_kinship(list(xb), scaled=False)), ref['binary_unscaled'])
    assert np.array_equal(np.asarray(kinship.calc_ibs_kinship(xd, 'diploid_int', scaled=False)), ref['diploid_unscaled'])
    np.testing.assert_allclose(np.asarray(kinship.calc_ibs_kinship(list(xb))), ref['binary_scaled'], rtol=1e-14)
    np.testing.assert_allclose(np.asarray(kinship.calc_ibs_kinship(xd, 'diploid_int')), ref['diploid_scaled'], rtol=1e-14)
    # IBD: the reference accumulates in float32 (kinship.py:62,69)
    np.testing.assert_allclose(np.asarray(kinship.calc_ibd_kinship(list(xd))), ref['ibd_scaled'], rtol=2e-4, atol=2e-6)
    np.testing.assert_allclose(np.asarray(kinship.calc_ibd_kinship(list(xd), scaled=False)), ref['ibd_unscaled'], rtol=2e-4, atol=2e-6)


@pytest.mark.parametrize('impl', ['dmma', 'tcgen05'])
def test_emmax_ft10_vs_reference_run(ctx, impl):
    from mixmogam_b200 import linear_models as lm
    ref = golden('ref_emmax_ft10_n198.npz')
    e = golden('emmax_ft10_n198.npz')
    snps, y, K = e['snps'], e['y'], e['K']
    lmm = lm.LinearMixedModel(y, scan_impl=impl)
    lmm.add_random_effect(K)
    r = lmm.emmax_f_test(list(snps), emma_num=0)
    assert_ps_close(r, ref)
    assert_reml_consistent(lmm, r, ref)
    scan_at_reference_delta(lmm, snps, ref)
    rb = scan_at_reference_delta(lmm, snps, ref, 'wb_', with_betas=True)
    b, bref = np.asarray(rb['betas'], dtype=np.float64), ref['wb_betas']
    np.testing.assert_allclose(b[:, -1], bref[:, -1], rtol=2e-3, atol=2e-5)          # the SNP effect
    re = lm.emmax(snps[:400], y, K, emma_num=5, scan_impl=impl)
    assert_ps_close(re, ref, 'emma5_', top=10)


def test_get_reml_vs_reference_run(ctx):
    from mixmogam_b200 import linear_models as lm
    for name, ref_name in (('emmax_ft10_n198.npz', 'ref_emmax_ft10_n198.npz'),
                           ('emmax_diploid_n400.npz', 'ref_emmax_diploid_n400.npz')):
        ref = golden(ref_name)
        e = golden(name)
        lmm = lm.LinearMixedModel(e['y'])
        lmm.add_random_effect(e['K'])
        res = lmm.get_REML()
        eig_R = lmm._last_reml['eig_R']
        ev, sq = np.asarray(eig_R['values']), np.asarray(lmm._etas(eig_R, lmm.Y)) ** 2
        assert lmm._rell_(res['delta'], ev, sq) >= lmm._rell_(float(ref['reml_delta']), ev, sq) - 1e-6
        if name.startswith('emmax_diploid'):            # curved likelihood, interior optimum: the numbers agree
            for k in ('delta', 'max_ll', 'vg', 've'):
                np.testing.assert_allclose(float(res[k]), float(ref['reml_' + k]), rtol=2e-4)


def test_snp_priors_vs_reference_run(ctx):
    from mixmogam_b200 import linear_models as lm
    ref = golden('ref_emmax_ft10_n198.npz')
    e = golden('emmax_ft10_n198.npz')
    lmm = lm.LinearMixedModel(e['y'])
    lmm.add_random_effect(e['K'])
    r = scan_at_reference_delta(lmm, e['snps'][:300], ref, 'priors_', snp_priors=ref['priors'])
    for k in ('bfs', 'pos', 'ppas'):
        np.testing.assert_allclose(np.asarray(r[k], dtype=np.float64), ref['priors_' + k], rtol=5e-3, atol=1e-12)


@pytest.mark.parametrize('impl', ['dmma', 'tcgen05'])
def test_emmax_diploid_cofactor_Z_vs_reference_run(ctx, impl):
    from mixmogam_b200 import linear_models as lm
    ref = golden('ref_emmax_diploid_n400.npz')
    e = golden('emmax_diploid_n400.npz')
    snps, y, K, cof = e['snps'], e['y'], e['K'], e['cofactor']
    lmm = lm.LinearMixedModel(y, scan_impl=impl)
    lmm.add_random_effect(K)
    r = lmm.emmax_f_test(snps, emma_num=0)
    assert_ps_close(r, ref)
    assert_reml_consistent(lmm, r, ref)
    assert abs(float(r['pseudo_heritability']) - float(ref['pseudo_heritability'])) < 1e-5    # interior optimum, refined in both
    scan_at_reference_delta(lmm, snps, ref)
    # SNP 17 IS the cofactor: its rotated genotype is ~0 after the projection and its p-value is float32 noise in
    # the reference; compare everything else
    ok = np.arange(len(snps)) != 17
    lmc = lm.LinearMixedModel(y, scan_impl=impl)
    lmc.add_random_effect(K)
    lmc.add_factor(cof)
    rc = lmc.emmax_f_test(snps, emma_num=0)
    assert_ps_close(rc, ref, 'cof_', skip=ok)
    assert_reml_consistent(lmc, rc, ref, 'cof_')
    scan_at_reference_delta(lmc, snps, ref, 'cof_', skip=ok)
    # replicate design (Z, linear_models.py:1795-1800,1296-1297)
    s = snps[:800, :100][ref['z_keep']]
    Z = ref['Z']
    lmz = lm.LinearMixedModel(ref['yz'], scan_impl=impl)
    lmz.add_random_effect(Z @ np.asarray(K)[:100, :100] @ Z.T)
    rz = lmz.emmax_f_test(s, Z=Z, emma_num=0)
    assert_ps_close(rz, ref, 'z_')
    assert_reml_consistent(lmz, rz, ref, 'z_')
    scan_at_reference_delta(lmz, s, ref, 'z_', Z=Z)


def test_permutations_vs_reference_run(ctx):
    from mixmogam_b200 import linear_models as lm
    ref = golden('ref_perm_n120.npz')
    e = golden('perm_n120.npz')
    lmm = lm.LinearMixedModel(e['y'])
    lmm.add_random_effect(e['K'])
    np.random.seed(int(ref['seed']))
    pr = lmm._emmax_permutations_(e['snps'], e['K'], ref['H_sqrt_inv'].astype(np.float64), num_perm=25)
    np.testing.assert_allclose(-np.log10(pr['min_ps']), -np.log10(ref['min_ps']), atol=1e-3)
    np.testing.assert_allclose(pr['max_f_stats'], ref['max_f_stats'], rtol=1e-3)
    np.testing.assert_allclose(np.asarray(lmm.Y).reshape(-1), ref['Y_after'], atol=1e-6)


def test_hdf5_entry_points_vs_reference_run(ctx):
    from mixmogam_b200 import hdf5_data
    ref = golden('ref_hdf5_n198.npz')
    snps = golden('ibs_diploid_n198.npz')['snps']
    chroms = [snps[:1700], snps[1700:]]
    gg = {}
    for i, x in enumerate(chroms):
        gg['chrom_%d' % (i + 1)] = {'raw_snps': x, 'freqs': x.mean(1) / 2.0, 'positions': np.arange(len(x)) * 100 + 1}
    f = {'genot_data': gg, 'indiv_data': {'indiv_ids': np.arange(198), 'phenotypes': ref['y']},
         'num_snps': np.array(len(snps))}
    out = {}
    hdf5_data.run_emmax(f, out, min_maf=0.1)
    assert int(out['num_snps']) == int(ref['num_snps'])
    # here the reference's float32 secant does not meet tol=1e-6 under NEP-50 numpy and it keeps the grid optimum
    # (delta = 2.2255 against the refined 2.1097): p-values stay within a few 1e-2, the ranking is the same
    for c in ('chrom_1', 'chrom_2'):
        a, b = -np.log10(out['chrom_results'][c]['ps']), -np.log10(ref[c + '_ps'])
        assert np.max(np.abs(a - b)) <= 5e-2
        assert np.array_equal(np.argsort(-a, kind='stable')[:10], np.argsort(-b, kind='stable')[:10])
        assert np.array_equal(out['chrom_results'][c]['positions'], ref[c + '_positions'])
    assert abs(float(out['pseudo_heritability']) - 0.32157013467222023) < 1e-6      # oracle float64 / numpy-1 float32 value
    assert abs(float(out['pseudo_heritability']) - float(ref['pseudo_heritability'])) < 0.02
    outp = {}
    np.random.seed(3)
    hdf5_data.run_emmax_perm(f, outp, min_maf=0.1, num_perm=40)
    np.testing.assert_allclose(outp['kinship'], ref['perm_kinship'], rtol=2e-4, atol=2e-6)
    assert int(outp['num_snps']) == int(ref['perm_num_snps'])
    # the permuted phenotypes are shuffles of the ROTATED residual: eigenvector signs differ between cuSOLVER and
    # LAPACK, so the null distribution agrees in distribution only (40 draws): compare medians on the log scale
    a, b = np.median(-np.log10(outp['perm_min_ps'])), np.median(-np.log10(ref['perm_min_ps']))
    assert abs(a - b) < 0.5
    f2 = {'genot_data': gg, 'indiv_data': {'indiv_ids': np.arange(198), 'phenotypes': ref['y']}}
    hdf5_data.calculate_ibd_kinship(f2)
    np.testing.assert_allclose(f2['kinship'], ref['ibd_kinship_nofilter'], rtol=2e-4, atol=2e-6)


def test_fast_f_test_vs_oracle_and_reference_run(ctx):
    """LinearModel.fast_f_test (linear_models.py:196-257): the scan with M = I - QQ' and no kinship, against the
    float64 oracle at the north-star tolerance and against the reference's own float32 run; plain, with a cofactor
    (SNP 17 is collinear with it: excluded, the reference's value there is noise) and with_betas."""
    from mixmogam_b200 import linear_models as lm
    from oracle import reference_py3 as o
    ref = golden('ref_fast_f_test_n400.npz')
    e = golden('emmax_diploid_n400.npz')
    snps, y, cof = e['snps'], e['y'], e['cofactor']
    ctx.invalidate_snps()
    r = lm.LinearModel(y).fast_f_test(snps)
    ro = o.LinearModel(list(y), dtype='double').fast_f_test(list(snps))
    assert neglog10_rel_err(r['ps'], ro['ps']) < 1e-6
    np.testing.assert_allclose(r['rss'], ro['rss'], rtol=1e-9)
    np.testing.assert_allclose(np.asarray(r['h0_rss']).reshape(-1), np.asarray(ro['h0_rss']).reshape(-1), rtol=1e-12)
    assert np.max(np.abs(np.log10(r['ps']) - np.log10(ref['ps']))) < 1e-2
    assert np.array_equal(np.argsort(r['ps'], kind='stable')[:20], np.argsort(ref['ps'], kind='stable')[:20])
    ok = np.arange(800) != 17
    m = lm.LinearModel(y)
    m.add_factor(cof)
    r2 = m.fast_f_test(snps[:800])
    mo = o.LinearModel(list(y), dtype='double')
    mo.add_factor(cof)
    ro2 = mo.fast_f_test(list(snps[:800]))
    assert neglog10_rel_err(r2['ps'][ok], ro2['ps'][ok]) < 1e-6
    assert np.max(np.abs(np.log10(r2['ps'][ok]) - np.log10(ref['cof_ps'][ok]))) < 1e-2
    np.testing.assert_allclose(r2['h0_betas'], ro2['h0_betas'], rtol=1e-9, atol=1e-12)
    ok3 = np.arange(300) != 17
    r3 = m.fast_f_test(snps[:300], with_betas=True)
    ro3 = mo.fast_f_test(list(snps[:300]), with_betas=True)
    assert neglog10_rel_err(r3['ps'][ok3], ro3['ps'][ok3]) < 1e-6
    b_gpu = np.array([b for b, k in zip(r3['betas'], ok3) if k])          # the collinear SNP keeps the 2 null betas (ragged list)
    b_ora = np.array([b for b, k in zip(ro3['betas'], ok3) if k])
    np.testing.assert_allclose(b_gpu, b_ora, rtol=1e-6, atol=1e-9)
    ctx.invalidate_snps()
