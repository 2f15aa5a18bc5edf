"""
Generates the committed golden fixtures in tests/golden/ from the oracle
(oracle/reference_py3.py) with fixed seeds.  Run in the build container:

    python tests/golden/make_golden.py

The FT10 phenotype (phenotype_id 5, 198 accessions) is read from the
reference's data file /root/reference/at_data/199_phenotypes.csv when present
(config 1 of BASELINE.json); the GPU box never reads /root/reference, it reads
the .npz written here.
"""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import reference_py3 as o  # noqa: E402

warnings.simplefilter('ignore')


def save(name, **kw):
    path = os.path.join(HERE, name)
    np.savez_compressed(path, **kw)
    print('%-28s %8.1f KB' % (name, os.path.getsize(path) / 1024.0))


def ft10():
    p = '/root/reference/at_data/199_phenotypes.csv'
    import pandas as pd
    d = pd.read_csv(p)
    g = d[d.phenotype_id == 5].sort_values('ecotype_id')
    return g.ecotype_id.to_numpy(np.int64), g.value.to_numpy(np.float64)


def scan_outputs(prefix, r):
    out = {}
    for k in ('ps', 'f_stats', 'rss', 'var_perc'):
        out[prefix + k] = np.asarray(r[k], dtype=np.float64)
    out[prefix + 'h0_rss'] = np.asarray(r['h0_rss'], dtype=np.float64).reshape(-1)
    out[prefix + 'h0_betas'] = np.asarray(r['h0_betas'], dtype=np.float64)
    for k in ('pseudo_heritability', 've', 'vg', 'max_ll', '_delta'):
        out[prefix + k.lstrip('_')] = np.float64(r[k])
    if 'betas' in r:
        out[prefix + 'betas'] = np.asarray(r['betas'], dtype=np.float64)
    return out


def main():
    # ---- kinship, both codings, literal loops --------------------------------
    xb = o.synth_genotypes(700, 37, 'binary', seed=20240601)
    save('ibs_binary_n37.npz', snps=xb,
         K_unscaled=o.calc_ibs_kinship(list(xb), 'binary', scaled=False),
         K_scaled=o.calc_ibs_kinship(list(xb), 'binary', scaled=True))
    xd = o.synth_genotypes(700, 37, 'diploid_int', seed=20240602)
    save('ibs_diploid_n37.npz', snps=xd,
         K_unscaled=o.calc_ibs_kinship(list(xd), 'diploid_int', scaled=False),
         K_scaled=o.calc_ibs_kinship(list(xd), 'diploid_int', scaled=True))
    xd2 = o.synth_genotypes(3000, 198, 'diploid_int', seed=20240603)
    save('ibs_diploid_n198.npz', snps=xd2,
         K_unscaled=o.calc_ibs_kinship_diploid_fast(xd2, scaled=False),
         K_scaled=o.calc_ibs_kinship_diploid_fast(xd2, scaled=True))
    save('ibd_n37.npz', snps=xd,
         K_single=o.calc_ibd_kinship(list(xd), dtype='single'),
         K_double=o.calc_ibd_kinship(list(xd), dtype='double'),
         K_double_unscaled=o.calc_ibd_kinship(list(xd), dtype='double', scaled=False))
    chroms = [xd2[:1700], xd2[1700:]]
    freqs = [c.mean(1) / 2.0 for c in chroms]
    kh, nh = o.hdf5_ibd_kinship(chroms, freqs, min_maf=0.1, chunk_size=1000, dtype='double')
    save('ibd_hdf5_n198.npz', freqs0=freqs[0], freqs1=freqs[1], K=kh, n_snps=np.int64(nh))

    # ---- EMMAX config 1 stand-in: FT10 on 198 accessions, binary genotypes -----
    ids, y = ft10()
    x1 = o.synth_genotypes(3000, 198, 'binary', seed=20240604)
    K1 = o.calc_ibs_kinship(list(x1), 'binary')
    out = dict(snps=x1, y=y, ecotype_ids=ids, K=K1)
    for dt in ('single', 'double'):
        r = o.emmax(list(x1), y, K1, dtype=dt)
        out.update(scan_outputs(dt + '_', r))
    rb = o.emmax(list(x1), y, K1, with_betas=True, dtype='double')
    out.update(scan_outputs('double_wb_', rb))
    lmm = o.LinearMixedModel(y, 'double')
    lmm.add_random_effect(K1)
    res = lmm.get_REML()
    out.update(reml_delta=np.float64(res['delta']), reml_max_ll=np.float64(res['max_ll']),
               reml_lls=np.asarray(res['_lls']), reml_dlls=np.asarray(res['_dlls']),
               reml_vg=np.float64(res['vg']), reml_ve=np.float64(res['ve']),
               reml_HtH=(res['H_sqrt_inv'].T @ res['H_sqrt_inv']),
               reml_beta=np.asarray(res['beta']).reshape(-1),
               reml_mahalanobis_rss=np.asarray(res['mahalanobis_rss']).reshape(-1),
               reml_eigL_values=np.asarray(res['eig_L']['values']),
               reml_eigR_values=np.asarray(res['_eig_R']['values']))
    re = o.emmax(list(x1[:400]), y, K1, emma_num=5, dtype='double')
    out.update(scan_outputs('double_emma5_', re))
    save('emmax_ft10_n198.npz', **out)

    # ---- EMMAX diploid, n=400, one cofactor -------------------------------------
    x2 = o.synth_genotypes(2500, 400, 'diploid_int', seed=20240605)
    K2 = o.calc_ibs_kinship_diploid_fast(x2)
    y2 = o.synth_phenotype(x2, K2, seed=11)
    cof = x2[17].astype(np.float64)
    out = dict(snps=x2, y=y2, K=K2, cofactor=cof)
    for dt in ('single', 'double'):
        r = o.emmax(list(x2), y2, K2, dtype=dt)
        out.update(scan_outputs(dt + '_', r))
    rc = o.emmax(list(x2), y2, K2, cofactors=[cof], dtype='double')
    out.update(scan_outputs('double_cof_', rc))
    save('emmax_diploid_n400.npz', **out)

    # ---- permutations (quirks of linear_models.py:1125-1175) -------------------
    x3 = o.synth_genotypes(600, 120, 'diploid_int', seed=20240606)
    K3 = o.calc_ibs_kinship_diploid_fast(x3)
    y3 = o.synth_phenotype(x3, K3, seed=13)
    lmm = o.LinearMixedModel(y3, 'double')
    lmm.add_random_effect(K3)
    res = lmm.get_REML()
    np.random.seed(20240607)
    pr = lmm._emmax_permutations_(x3.astype(np.float64), K3, res['H_sqrt_inv'], num_perm=25)
    save('perm_n120.npz', snps=x3, y=y3, K=K3, seed=np.int64(20240607), Ys=pr['_Ys'], H_sqrt_inv=res['H_sqrt_inv'],
         min_ps=pr['min_ps'], max_f_stats=pr['max_f_stats'], h0_rss=np.asarray(pr['_h0_rss']).reshape(-1),
         delta=np.float64(res['delta']))

    # ---- F survival function known answers (scipy.stats.f.sf) ------------------
    from scipy import stats
    f = np.concatenate([[0.0, 1e-12, 1e-6, 1e-3, 0.1, 0.5, 1.0, 2.0, 3.84, 10.0, 30.0, 100.0, 300.0,
                         1e3, 3e3, 1e4, 1e5, 1e6], np.exp(np.linspace(-8, 9, 60))])
    dfd = np.array([3.0, 35.0, 196.0, 398.0, 1998.0, 9998.0, 49998.0])
    F, D = np.meshgrid(f, dfd, indexing='ij')
    save('f_sf.npz', f=F, dfd=D, sf=stats.f.sf(F, 1, D),
         f2=F, sf2=stats.f.sf(F, 2, D), sf3=stats.f.sf(F, 3, D))


if __name__ == '__main__':
    main()
