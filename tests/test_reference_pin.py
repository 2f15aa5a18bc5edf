"""CPU: the oracle restatement is pinned to the REFERENCE'S OWN CODE.

tests/golden/ref_*.npz hold outputs of the unmodified reference sources executed in the build container
(tests/golden/py2shim.py + make_reference_golden.py).  The oracle in dtype='single', promotion='numpy2' mode
(same interpreter, same numpy/scipy/LAPACK as that run) must reproduce them BIT FOR BIT; the float64 mode the
CUDA path is compared with is the same code with the dtype switched, and must stay inside the reference's
float32 noise.
"""
import warnings

import numpy as np
import pytest

from conftest import golden
from oracle import reference_py3 as o

warnings.simplefilter('ignore')

SCAN_KEYS = ('ps', 'f_stats', 'rss', 'var_perc', 'h0_rss', 'h0_betas')
REML_KEYS = ('pseudo_heritability', 've', 'vg', 'max_ll')


def assert_scan_bit_exact(r, ref, prefix, keys=SCAN_KEYS + REML_KEYS):
    for k in keys:
        a = np.asarray(r[k], dtype=np.float64).reshape(-1)
        b = np.asarray(ref[prefix + k], dtype=np.float64).reshape(-1)
        assert np.array_equal(a, b), '%s%s differs: max abs %g' % (prefix, k, np.max(np.abs(a - b)))


def test_kinship_bit_exact_vs_reference_run():
    ref = golden('ref_kinship_n37.npz')
    xb = golden('ibs_binary_n37.npz')['snps']
    xd = golden('ibs_diploid_n37.npz')['snps']
    assert np.array_equal(o.calc_ibs_kinship(list(xb), 'binary', scaled=False), ref['binary_unscaled'])
    assert np.array_equal(o.calc_ibs_kinship(list(xb), 'binary'), ref['binary_scaled'])
    assert np.array_equal(ref['binary_chunk64'], ref['binary_unscaled'])
    assert np.array_equal(o.calc_ibs_kinship(list(xd), 'diploid_int', scaled=False), ref['diploid_unscaled'])
    assert np.array_equal(o.calc_ibs_kinship(list(xd), 'diploid_int'), ref['diploid_scaled'])
    assert np.array_equal(o.calc_ibs_kinship_diploid_fast(xd, scaled=False), ref['diploid_unscaled'])
    assert np.array_equal(o.calc_ibd_kinship(list(xd), dtype='single'), ref['ibd_scaled'])
    assert np.array_equal(o.calc_ibd_kinship(list(xd), dtype='single', scaled=False), ref['ibd_unscaled'])
    # the committed oracle fixtures are therefore the reference's own numbers
    assert np.array_equal(golden('ibs_binary_n37.npz')['K_unscaled'], ref['binary_unscaled'])
    assert np.array_equal(golden('ibs_diploid_n37.npz')['K_unscaled'], ref['diploid_unscaled'])


def test_emmax_ft10_bit_exact_vs_reference_run():
    ref = golden('ref_emmax_ft10_n198.npz')
    e = golden('emmax_ft10_n198.npz')
    snps, y, K = e['snps'], e['y'], e['K']
    assert_scan_bit_exact(o.emmax(list(snps), y, K, dtype='single', promotion='numpy2'), ref, '')
    rb = o.emmax(list(snps), y, K, with_betas=True, dtype='single', promotion='numpy2')
    assert_scan_bit_exact(rb, ref, 'wb_')
    assert np.array_equal(np.asarray(rb['betas'], dtype=np.float64), ref['wb_betas'])
    assert_scan_bit_exact(o.emmax(list(snps[:400]), y, K, emma_num=5, dtype='single', promotion='numpy2'), ref, 'emma5_')


def test_get_reml_bit_exact_vs_reference_run():
    ref = golden('ref_emmax_ft10_n198.npz')
    e = golden('emmax_ft10_n198.npz')
    lmm = o.LinearMixedModel(e['y'], 'single', promotion='numpy2')
    lmm.add_random_effect(e['K'])
    res = lmm.get_REML()
    for k in ('delta', 'max_ll', 'vg', 've', 'pseudo_heritability'):
        assert np.float64(res[k]) == ref['reml_' + k], k
    assert np.array_equal(np.asarray(res['beta'], dtype=np.float64).reshape(-1), ref['reml_beta'])
    assert np.array_equal(np.asarray(res['mahalanobis_rss'], dtype=np.float64).reshape(-1), ref['reml_mahalanobis_rss'])
    assert np.array_equal(np.asarray(res['eig_L']['values'], dtype=np.float64), ref['reml_eigL_values'])


def test_snp_priors_bit_exact_vs_reference_run():
    ref = golden('ref_emmax_ft10_n198.npz')
    e = golden('emmax_ft10_n198.npz')
    lmm = o.LinearMixedModel(e['y'], 'single', promotion='numpy2')
    lmm.add_random_effect(e['K'])
    r = lmm.emmax_f_test(list(e['snps'][:300]), snp_priors=ref['priors'], emma_num=0)
    assert_scan_bit_exact(r, ref, 'priors_', keys=SCAN_KEYS + ('bfs', 'pos', 'ppas'))


def test_emmax_diploid_cofactor_and_Z_bit_exact_vs_reference_run():
    ref = golden('ref_emmax_diploid_n400.npz')
    e = golden('emmax_diploid_n400.npz')
    snps, y, K, cof = e['snps'], e['y'], e['K'], e['cofactor']
    assert_scan_bit_exact(o.emmax(list(snps), y, K, dtype='single', promotion='numpy2'), ref, '')
    assert_scan_bit_exact(o.emmax(list(snps), y, K, cofactors=[cof], dtype='single', promotion='numpy2'), ref, 'cof_')
    lmm = o.LinearMixedModel(y, 'single', promotion='numpy2')
    lmm.add_random_effect(K)
    res = lmm.get_REML()
    for k in ('delta', 'max_ll', 'vg', 've'):
        assert np.float64(res[k]) == ref['reml_' + k], k
    assert 1e-3 < float(res['delta']) < 1e3            # an interior optimum: the secant refinement is exercised
    sub = snps[:800, :100][ref['z_keep']]
    rz = o.emmax(list(sub), ref['yz'], np.asarray(K)[:100, :100], Z=ref['Z'], dtype='single', promotion='numpy2')
    assert_scan_bit_exact(rz, ref, 'z_')


def test_fast_f_test_bit_exact_vs_reference_run():
    """LinearModel.fast_f_test (linear_models.py:196-257), the OLS sibling of the EMMAX scan: plain, with a cofactor, and
    with_betas (where SNP 17 is collinear with the cofactor-augmented design and the reference stores a NaN residue)."""
    ref = golden('ref_fast_f_test_n400.npz')
    e = golden('emmax_diploid_n400.npz')
    snps, y, cof = e['snps'], e['y'], e['cofactor']
    keys = ('ps', 'f_stats', 'rss', 'var_perc', 'h0_rss', 'h0_betas')

    def same(r, prefix, extra=()):
        for k in keys + tuple(extra):
            a = np.asarray(r[k], dtype=np.float64).reshape(-1)
            b = np.asarray(ref[prefix + k], dtype=np.float64).reshape(-1)
            assert np.array_equal(a, b, equal_nan=True), prefix + k

    same(o.LinearModel(list(y)).fast_f_test(list(snps)), '')
    m = o.LinearModel(list(y))
    m.add_factor(cof)
    same(m.fast_f_test(list(snps[:800])), 'cof_')
    m = o.LinearModel(list(y))
    m.add_factor(cof)
    same(m.fast_f_test(list(snps[:300]), with_betas=True), 'wb_', extra=('betas',))
    # the float64 mode the GPU path is held to stays within the reference's float32 noise
    rd = o.LinearModel(list(y), dtype='double').fast_f_test(list(snps))
    assert np.max(np.abs(np.log10(rd['ps']) - np.log10(ref['ps']))) < 1e-2


def test_permutations_bit_exact_vs_reference_run():
    ref = golden('ref_perm_n120.npz')
    e = golden('perm_n120.npz')
    lmm = o.LinearMixedModel(e['y'], 'single', promotion='numpy2')
    lmm.add_random_effect(e['K'])
    res = lmm.get_REML()
    assert np.array_equal(np.asarray(res['H_sqrt_inv']), ref['H_sqrt_inv'])
    np.random.seed(int(ref['seed']))
    pr = lmm._emmax_permutations_(e['snps'].astype(np.float64), e['K'], res['H_sqrt_inv'], num_perm=25)
    assert np.array_equal(np.asarray(pr['min_ps'], dtype=np.float64), ref['min_ps'])
    assert np.array_equal(np.asarray(pr['max_f_stats'], dtype=np.float64), ref['max_f_stats'])
    assert np.array_equal(np.asarray(lmm.Y, dtype=np.float64).reshape(-1), ref['Y_after'])     # :1140 mutates the model


def test_hdf5_paths_bit_exact_vs_reference_run():
    ref = golden('ref_hdf5_n198.npz')
    snps = golden('ibs_diploid_n198.npz')['snps']
    chroms = [snps[:1700], snps[1700:]]
    freqs = [c.mean(1) / 2.0 for c in chroms]
    k, n_snps = o.hdf5_ibd_kinship(chroms, freqs, min_maf=0.1, chunk_size=1000, dtype='single')
    assert np.array_equal(np.asarray(k, dtype=np.float64), ref['perm_kinship'])
    k0, _ = o.hdf5_ibd_kinship(chroms, None, min_maf=None, chunk_size=1000, dtype='single')
    assert np.array_equal(np.asarray(k0, dtype=np.float64), ref['ibd_kinship_nofilter'])
    lmm = o.LinearMixedModel(ref['y'], 'single', promotion='numpy2')
    lmm.add_random_effect(k)
    eig_L = lmm._get_eigen_L_()
    eig_R = lmm._get_eigen_R_(X=lmm.X)
    res = lmm.get_estimates(eig_L, method='REML', eig_R=eig_R)
    for key in REML_KEYS:
        assert np.float64(res[key]) == ref[key], key
    for ci, c in enumerate(chroms):
        keep = np.minimum(freqs[ci], 1 - freqs[ci]) > 0.1
        r = lmm._emmax_f_test_(c[keep], res['H_sqrt_inv'], with_betas=False, emma_num=0, eig_L=eig_L)
        assert np.array_equal(np.asarray(r['ps'], dtype=np.float64), ref['chrom_%d_ps' % (ci + 1)])
        assert np.array_equal((np.arange(len(c)) * 100 + 1)[keep], ref['chrom_%d_positions' % (ci + 1)])
    # the reference stores the LAST chromosome's post-filter count as num_snps in the permutation output (:253,:289)
    keep_last = np.minimum(freqs[-1], 1 - freqs[-1]) > 0.1
    assert int(ref['perm_num_snps']) == int(keep_last.sum())


@pytest.mark.parametrize('name,ref_name', [('emmax_ft10_n198.npz', 'ref_emmax_ft10_n198.npz'),
                                           ('emmax_diploid_n400.npz', 'ref_emmax_diploid_n400.npz')])
def test_float64_oracle_within_reference_float32_noise(name, ref_name):
    """The float64 mode (what the CUDA path is held to at 1e-6) against the reference's own float32 run."""
    ref = golden(ref_name)
    e = golden(name)
    d = np.abs(np.log10(e['double_ps']) - np.log10(ref['ps']))
    assert d.max() < 1e-2
    assert np.array_equal(np.argsort(e['double_ps'])[:20], np.argsort(ref['ps'])[:20])
    assert abs(float(e['double_pseudo_heritability']) - float(ref['pseudo_heritability'])) < 1e-4


def test_ml_emma_gxt_bit_exact_vs_reference_run():
    """get_ML (linear_models.py:672-696 with the ML branch :811-824), expedited_REML_t_test (:931-968) and
    emmax_w_two_env -> _emmax_GxT_f_test_ (:1749-1787, :1422-1514), oracle vs the reference's own code, bit for bit; the
    float64 mode the GPU path is held to stays inside the reference's float32 noise."""
    ref = golden('ref_ml_emma_gxt_n400.npz')
    e = golden('emmax_diploid_n400.npz')
    snps, y, K, cof = e['snps'], e['y'], e['K'], e['cofactor']

    def same(a, key):
        a = np.asarray(a, dtype=np.float64).reshape(-1)
        b = np.asarray(ref[key], dtype=np.float64).reshape(-1)
        assert np.array_equal(a, b, equal_nan=True), '%s differs: max abs %g' % (key, np.max(np.abs(a - b)))

    for tag, cofs in (('', []), ('cof_', [cof])):
        lmm = o.LinearMixedModel(list(y), 'single', promotion='numpy2')
        lmm.add_random_effect(K)
        for c in cofs:
            lmm.add_factor(c)
        r = lmm.get_ML()
        for k in ('delta', 'max_ll', 'vg', 've', 'pseudo_heritability', 'beta', 'mahalanobis_rss', 'rss'):
            same(r[k], 'ml_' + tag + k)
        rr = lmm.expedited_REML_t_test(list(snps[:12]))
        for k in ('ps', 'f_stats', 'vgs', 'ves', 'var_perc', 'max_lls', 'rss', 'betas'):
            same(rr[k], 'emma_' + tag + k)
        rd = o.LinearMixedModel(list(y), 'double')
        rd.add_random_effect(K)
        for c in cofs:
            rd.add_factor(c)
        md = rd.get_ML()
        # (the float32 likelihood of the reference is flat to its own rounding around the optimum: delta moves by a few per cent)
        assert abs(md['delta'] / float(ref['ml_' + tag + 'delta']) - 1) < 0.1 and abs(md['max_ll'] - float(ref['ml_' + tag + 'max_ll'])) < 5e-2
    E, ye = ref['E'], ref['ye']
    for tag, cofs in (('', None), ('cof_', [cof])):
        r = o.emmax_w_two_env(list(snps[:600]), list(ye), K, E, cofs, dtype='single', promotion='numpy2')
        for part in ('g_res', 'gt_res', 'gt_g_res'):
            for k in ('ps', 'f_stats', 'var_perc'):
                same(r[part][k], 'gxt_%s%s_%s' % (tag, part, k))
        for part in ('g_res', 'gt_res'):
            same(r[part]['rss'], 'gxt_%s%s_rss' % (tag, part))
            same(r[part]['betas'], 'gxt_%s%s_betas' % (tag, part))
        same(r['g_res']['h0_rss'], 'gxt_%sh0_rss' % tag)
        for k in ('pseudo_heritability', 've', 'vg', 'max_ll'):
            same(r[k], 'gxt_%s%s' % (tag, k))
        rd = o.emmax_w_two_env(list(snps[:600]), list(ye), K, E, cofs, dtype='double')
        ok = ~np.all(snps[:600] == cof[None, :], axis=1) if cofs else np.ones(600, dtype=bool)     # the cofactor itself: collinear, noise
        for part in ('g_res', 'gt_res', 'gt_g_res'):
            d = np.abs(np.log10(rd[part]['ps']) - np.log10(ref['gxt_%s%s_ps' % (tag, part)]))
            assert np.max(d[ok]) < 5e-2                                   # float32 three-column lstsq
