"""
Golden vectors produced by the REFERENCE'S OWN CODE, executed in the build container through
tests/golden/py2shim.py (the unmodified /root/reference sources, token-translated Python 2 -> 3 in memory,
`scipy` re-pointed at numpy for the removed aliases).  Run once here:

    python tests/golden/make_reference_golden.py

Writes tests/golden/ref_*.npz (outputs only; the inputs are the ones already committed in the oracle
fixtures of make_golden.py, same seeds).  These files are the pin of the parity chain:

    reference source run here  ==  oracle(dtype='single', promotion='numpy2')      bit for bit   (CPU tests)
    oracle(dtype='double')     ~=  CUDA path       1e-6 relative in -log10 p                     (GPU tests)
    reference source run here  ~=  CUDA path       float32 noise of the reference: |d(-log10 p)| <= 1e-2,
                                                   identical top-20 ranking; kinship bit-exact    (GPU tests)

The GPU box has no /root/reference: tests only read the .npz written here.
"""
import contextlib
import io
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import py2shim  # noqa: E402

warnings.simplefilter('ignore')


def quiet(f, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return f(*a, **k)


def save(name, **kw):
    path = os.path.join(HERE, name)
    np.savez_compressed(path, **kw)
    print('%-28s %8.1f KB' % (name, os.path.getsize(path) / 1024.0))


def g(name):
    return np.load(os.path.join(HERE, name))


def scan_outputs(prefix, r):
    out = {}
    for k in ('ps', 'f_stats', 'rss', 'var_perc'):
        out[prefix + k] = np.asarray(r[k], dtype=np.float64)
    out[prefix + 'h0_rss'] = np.asarray(r['h0_rss'], dtype=np.float64).reshape(-1)
    out[prefix + 'h0_betas'] = np.asarray(r['h0_betas'], dtype=np.float64)
    for k in ('pseudo_heritability', 've', 'vg', 'max_ll'):
        if k in r:
            out[prefix + k] = np.float64(r[k])
    if 'betas' in r:
        out[prefix + 'betas'] = np.asarray(r['betas'], dtype=np.float64)
    for k in ('bfs', 'pos', 'ppas'):
        if k in r:
            out[prefix + k] = np.asarray(r[k], dtype=np.float64)
    return out


def h5_file(chroms, y):
    f = py2shim.FakeGroup()
    gg = f.create_group('genot_data')
    for i, x in enumerate(chroms):
        cg = gg.create_group('chrom_%d' % (i + 1))
        cg.create_dataset('raw_snps', data=x)
        cg.create_dataset('freqs', data=x.mean(1) / 2.0)
        cg.create_dataset('positions', data=np.arange(len(x)) * 100 + 1)
    ig = f.create_group('indiv_data')
    ig.create_dataset('indiv_ids', data=np.arange(chroms[0].shape[1]))
    ig.create_dataset('phenotypes', data=y)
    f.create_dataset('num_snps', data=np.array(sum(len(x) for x in chroms)))
    return f


def fast_f_test_golden(lm):
    """LinearModel.fast_f_test (linear_models.py:196-257): the OLS sibling of the EMMAX scan."""
    e = g('emmax_diploid_n400.npz')
    snps, y, cof = e['snps'], e['y'], e['cofactor']
    out = {}
    out.update(scan_outputs('', quiet(lm.LinearModel(list(y)).fast_f_test, list(snps))))
    m = lm.LinearModel(list(y))
    m.add_factor(cof)
    out.update(scan_outputs('cof_', quiet(m.fast_f_test, list(snps[:800]))))
    m = lm.LinearModel(list(y))
    m.add_factor(cof)
    out.update(scan_outputs('wb_', quiet(m.fast_f_test, list(snps[:300]), with_betas=True)))
    save('ref_fast_f_test_n400.npz', **out)


def ml_emma_gxt_golden(lm):
    """get_ML and the ML branch of get_estimates (linear_models.py:672-696, 811-824), the exact-EMMA
    refinement expedited_REML_t_test (:931-968, one eigendecomposition of S(K+I)S per SNP in the reference) and the G x E scan
    emmax_w_two_env -> _emmax_GxT_f_test_ (:1749-1787, :1422-1514)."""
    e = g('emmax_diploid_n400.npz')
    snps, y, K, cof = e['snps'], e['y'], e['K'], e['cofactor']
    out = {}
    for tag, cofs in (('', []), ('cof_', [cof])):
        lmm = lm.LinearMixedModel(list(y))
        lmm.add_random_effect(K)
        for c in cofs:
            lmm.add_factor(c)
        r = quiet(lmm.get_ML)
        for k in ('delta', 'max_ll', 'vg', 've', 'pseudo_heritability'):
            out['ml_' + tag + k] = np.float64(r[k])
        out['ml_' + tag + 'beta'] = np.asarray(r['beta'], dtype=np.float64).reshape(-1)
        out['ml_' + tag + 'mahalanobis_rss'] = np.asarray(r['mahalanobis_rss'], dtype=np.float64).reshape(-1)
        out['ml_' + tag + 'rss'] = np.asarray(r['rss'], dtype=np.float64).reshape(-1)
        rr = quiet(lmm.expedited_REML_t_test, list(snps[:12]))
        for k in ('ps', 'f_stats', 'vgs', 'ves', 'var_perc', 'max_lls', 'rss'):
            out['emma_' + tag + k] = np.asarray(rr[k], dtype=np.float64)
        out['emma_' + tag + 'betas'] = np.asarray(rr['betas'], dtype=np.float64)
    rng = np.random.Generator(np.random.PCG64(20240612))
    E = (rng.random(400) < 0.5).astype(np.float64).reshape(-1, 1)
    ye = np.asarray(y) + 0.8 * E[:, 0] * (snps[20] - snps[20].mean()) + 0.3 * E[:, 0]
    out.update(E=E, ye=ye)
    for tag, cofs in (('', None), ('cof_', [cof])):
        r = quiet(lm.emmax_w_two_env, list(snps[:600]), list(ye), K, E, cofs)
        for part in ('g_res', 'gt_res', 'gt_g_res'):
            for k in ('ps', 'f_stats', 'var_perc'):
                out['gxt_%s%s_%s' % (tag, part, k)] = np.asarray(r[part][k], dtype=np.float64)
        for part in ('g_res', 'gt_res'):
            out['gxt_%s%s_rss' % (tag, part)] = np.asarray(r[part]['rss'], dtype=np.float64)
            out['gxt_%s%s_betas' % (tag, part)] = np.asarray(r[part]['betas'], dtype=np.float64)
        out['gxt_%sh0_rss' % tag] = np.asarray(r['g_res']['h0_rss'], dtype=np.float64).reshape(-1)
        for k in ('pseudo_heritability', 've', 'vg', 'max_ll'):
            out['gxt_%s%s' % (tag, k)] = np.float64(r[k])
    save('ref_ml_emma_gxt_n400.npz', **out)


def main():
    kin = py2shim.load('kinship')
    lm = py2shim.load('linear_models')
    if sys.argv[1:] == ['fast_f_test']:
        return fast_f_test_golden(lm)
    if sys.argv[1:] == ['ml_emma_gxt']:
        return ml_emma_gxt_golden(lm)

    # ---- kinship.py:14-100, all three estimators, literal loops -------------------------------------
    xb = g('ibs_binary_n37.npz')['snps']
    xd = g('ibs_diploid_n37.npz')['snps']
    save('ref_kinship_n37.npz',
         binary_unscaled=np.asarray(quiet(kin.calc_ibs_kinship, list(xb), scaled=False)),
         binary_scaled=np.asarray(quiet(kin.calc_ibs_kinship, list(xb))),
         binary_chunk64=np.asarray(quiet(kin.calc_ibs_kinship, list(xb), chunk_size=64, scaled=False)),
         diploid_unscaled=np.asarray(quiet(kin.calc_ibs_kinship, list(xd), 'diploid_int', scaled=False)),
         diploid_scaled=np.asarray(quiet(kin.calc_ibs_kinship, list(xd), 'diploid_int')),
         ibd_scaled=np.asarray(quiet(kin.calc_ibd_kinship, list(xd))),
         ibd_unscaled=np.asarray(quiet(kin.calc_ibd_kinship, list(xd), scaled=False)))

    # ---- linear_models.emmax on the config-1 stand-in (FT10, 198 accessions, binary genotypes) -------
    e = g('emmax_ft10_n198.npz')
    snps, y, K = e['snps'], e['y'], e['K']
    out = {}
    out.update(scan_outputs('', quiet(lm.emmax, list(snps), list(y), K)))
    out.update(scan_outputs('wb_', quiet(lm.emmax, list(snps), list(y), K, with_betas=True)))
    out.update(scan_outputs('emma5_', quiet(lm.emmax, list(snps[:400]), list(y), K, emma_num=5)))
    lmm = lm.LinearMixedModel(list(y))
    lmm.add_random_effect(K)
    res = quiet(lmm.get_REML)
    out.update(reml_delta=np.float64(res['delta']), reml_max_ll=np.float64(res['max_ll']),
               reml_vg=np.float64(res['vg']), reml_ve=np.float64(res['ve']),
               reml_pseudo_heritability=np.float64(res['pseudo_heritability']),
               reml_beta=np.asarray(res['beta'], dtype=np.float64).reshape(-1),
               reml_mahalanobis_rss=np.asarray(res['mahalanobis_rss'], dtype=np.float64).reshape(-1),
               reml_eigL_values=np.asarray(res['eig_L']['values'], dtype=np.float64))
    priors = np.linspace(0.001, 0.05, 300)
    lmm = lm.LinearMixedModel(list(y))
    lmm.add_random_effect(K)
    out.update(scan_outputs('priors_', quiet(lmm.emmax_f_test, list(snps[:300]), snp_priors=priors, emma_num=0)))
    out['priors'] = priors
    save('ref_emmax_ft10_n198.npz', **out)

    # ---- diploid genotypes, n=400, interior REML optimum, one cofactor, replicate design Z ------------
    e = g('emmax_diploid_n400.npz')
    snps, y, K, cof = e['snps'], e['y'], e['K'], e['cofactor']
    out = {}
    out.update(scan_outputs('', quiet(lm.emmax, list(snps), list(y), K)))
    out.update(scan_outputs('cof_', quiet(lm.emmax, list(snps), list(y), K, cofactors=[cof])))
    lmm = lm.LinearMixedModel(list(y))
    lmm.add_random_effect(K)
    res = quiet(lmm.get_REML)
    out.update(reml_delta=np.float64(res['delta']), reml_max_ll=np.float64(res['max_ll']),
               reml_vg=np.float64(res['vg']), reml_ve=np.float64(res['ve']))
    # Z: 100 lines, 160 observations (some lines replicated), linear_models.py:1795-1800,1296-1297
    rng = np.random.Generator(np.random.PCG64(20240610))
    lines = np.concatenate([np.arange(100), rng.integers(0, 100, size=60)])
    Z = np.zeros((160, 100))
    Z[np.arange(160), lines] = 1.0
    sub = snps[:800, :100]
    keep = (sub.min(1) != sub.max(1))
    sub = sub[keep]
    Kz = np.asarray(e['K'])[:100, :100]
    yz = rng.standard_normal(160) + Z @ (sub[5] * 0.6)
    out.update(Z=Z, z_keep=keep, yz=yz)
    out.update(scan_outputs('z_', quiet(lm.emmax, list(sub), list(yz), Kz, Z=np.asmatrix(Z))))
    save('ref_emmax_diploid_n400.npz', **out)

    # ---- _emmax_permutations_ (linear_models.py:1125-1175), seeded global RNG ---------------------------
    e = g('perm_n120.npz')
    snps, y, K = e['snps'], e['y'], e['K']
    lmm = lm.LinearMixedModel(list(y))
    lmm.add_random_effect(K)
    res = quiet(lmm.get_REML)
    H = np.array(res['H_sqrt_inv'])
    np.random.seed(20240607)
    pr = quiet(lmm._emmax_permutations_, snps.astype(np.float64), K, res['H_sqrt_inv'], num_perm=25)
    save('ref_perm_n120.npz', seed=np.int64(20240607), H_sqrt_inv=H, delta=np.float64(res['delta']),
         min_ps=np.asarray(pr['min_ps'], dtype=np.float64), max_f_stats=np.asarray(pr['max_f_stats'], dtype=np.float64),
         Y_after=np.asarray(lmm.Y, dtype=np.float64).reshape(-1))

    # ---- hdf5_data.py entry points on an in-memory file with the plink2hdf5 layout ----------------------
    snps = g('ibs_diploid_n198.npz')['snps']
    chroms = [snps[:1700], snps[1700:]]
    rng = np.random.Generator(np.random.PCG64(20240611))
    y = rng.standard_normal(198) + 0.5 * snps[40] - 0.4 * snps[2000]
    files = {'in': h5_file(chroms, y)}
    h5 = py2shim.fake_h5py(files)
    hd = py2shim.load('hdf5_data', h5py_module=h5)
    quiet(hd.run_emmax, 'in', 'out', min_maf=0.1)
    o = files['out']
    out = dict(y=y, pseudo_heritability=np.float64(o['pseudo_heritability'].data), ve=np.float64(o['ve'].data),
               vg=np.float64(o['vg'].data), max_ll=np.float64(o['max_ll'].data), num_snps=np.int64(o['num_snps'].data))
    for c in ('chrom_1', 'chrom_2'):
        out[c + '_ps'] = np.asarray(o['chrom_results'][c]['ps'].data, dtype=np.float64)
        out[c + '_positions'] = np.asarray(o['chrom_results'][c]['positions'].data)
    np.random.seed(3)
    quiet(hd.run_emmax_perm, 'in', 'outp', min_maf=0.1, num_perm=40)
    o = files['outp']
    out.update(perm_kinship=np.asarray(o['kinship'].data, dtype=np.float64),
               perm_min_ps=np.asarray(o['perm_min_ps'].data, dtype=np.float64),
               perm_max_f_stats=np.asarray(o['perm_max_f_stats'].data, dtype=np.float64),
               perm_num_snps=np.int64(o['num_snps'].data),
               five_perc_perm_min_ps=np.float64(o['five_perc_perm_min_ps'].data))
    files2 = {'in': h5_file(chroms, y)}
    hd2 = py2shim.load('hdf5_data', h5py_module=py2shim.fake_h5py(files2), fresh=True)
    quiet(hd2.calculate_ibd_kinship, 'in')
    out['ibd_kinship_nofilter'] = np.asarray(files2['in']['kinship'].data, dtype=np.float64)
    save('ref_hdf5_n198.npz', **out)
    fast_f_test_golden(lm)
    ml_emma_gxt_golden(lm)


if __name__ == '__main__':
    main()
