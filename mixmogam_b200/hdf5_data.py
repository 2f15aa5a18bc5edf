"""
Drop-in for the EMMAX entry points of mixmogam's `hdf5_data` module (reference hdf5_data.py):

    calculate_ibd_kinship(hdf5_filename, chunk_size=1000, overwrite=False)                       :17-66
    run_emmax(hdf5_filename, out_file, min_maf=0.1, recalculate_kinship=True, chunk_size=1000)   :70-187
    run_emmax_perm(hdf5_filename, out_file, min_maf=0.1, recalculate_kinship=True,
                   chunk_size=1000, num_perm=500)                                                :191-351

Input layout (written by the reference's plink2hdf5.py:27-28,57-59,111-118,226):
    genot_data/<chrom>/{raw_snps int8 (m_c x n), freqs, positions, ...},  indiv_data/{indiv_ids, phenotypes},
    num_snps, optional kinship.
`hdf5_filename` / `out_file` may be file names (opened with h5py, imported lazily -- it is not part of
this image) or any mapping with that layout (nested dicts of numpy arrays); results are written with
`create_dataset` when the target has it, else by item assignment.

The numerics (IBD kinship, eigen, REML, per-chromosome scan, permutations) go through the same device
stages as linear_models; file handling stays on the host.
"""
import numpy as np

from . import _lib
from . import linear_models as lm

__all__ = ['calculate_ibd_kinship', 'run_emmax', 'run_emmax_perm']


def _open(f, mode='r'):
    if isinstance(f, str):
        try:
            import h5py
        except ImportError:
            raise ImportError('h5py is required to open %r; pass an in-memory mapping with the same layout instead' % f)
        return h5py.File(f, mode), True
    return f, False


def _put(g, name, data):
    if hasattr(g, 'create_dataset'):
        g.create_dataset(name, data=data)
    else:
        g[name] = np.asarray(data)


def _group(g, name):
    if hasattr(g, 'create_group'):
        return g.create_group(name)
    g[name] = {}
    return g[name]


def _arr(x):
    return np.asarray(x[...]) if hasattr(x, 'shape') and not isinstance(x, np.ndarray) else np.asarray(x)


def _ibd_kinship_device(ctx, gg, n_indivs, min_maf=None, use_normalized=False):
    """hdf5_data.py:30-62 / :84-115 / :205-237: K = sum over chromosomes of Z'Z / n_snps, then the inline
    scale_k.  Returns (K DeviceMatrix, n_snps)."""
    K = ctx.matrix(n_indivs, n_indivs)
    n_snps = 0
    for chrom in gg.keys():
        cg = gg[chrom]
        snps = _arr(cg['raw_snps'])
        mask = None
        if min_maf is not None:
            freqs = _arr(cg['freqs'])
            mafs = np.minimum(freqs, 1 - freqs)
            mask = (mafs > min_maf)                          # :91-96
        m_c, n = ctx.ensure_snps(snps)
        try:
            n_snps += ctx.kinship_ibd_accumulate(K, 0, m_c, mask)
        except _lib.MmgError as e:
            if e.code == -7:
                raise FloatingPointError('monomorphic SNP on chromosome %s: its standardised genotype is NaN '
                                         '(the reference silently propagates NaN here)' % chrom)
            raise
    Kh_scale = 1.0 / float(n_snps)                           # :58 / :111
    ones = np.full(n_indivs, Kh_scale)
    ctx.scale_rows(K, ones)
    ctx.scale_k(K)                                           # :59-62 / :112-115
    return K, n_snps


def calculate_ibd_kinship(hdf5_filename='/home/bv25/data/Ls154/Ls154_12.hdf5',
                          chunk_size=1000, overwrite=False, ctx=None):
    """
    Calculates a kinship matrix and stores it in the HDF5 file (hdf5_data.py:17-66).
    """
    ctx = ctx or _lib.get_context()
    h5f, opened = _open(hdf5_filename, 'r+')
    n_indivs = len(_arr(h5f['indiv_data']['indiv_ids']))
    if overwrite or 'kinship' not in h5f.keys():
        if 'kinship' in h5f.keys():
            del h5f['kinship']
        K, n_snps = _ibd_kinship_device(ctx, h5f['genot_data'], n_indivs, min_maf=None)
        k = K.download()
        K.free()
        _put(h5f, 'kinship', k)
    if opened:
        h5f.close()


def _fit_null(ctx, k, phenotypes):
    """hdf5_data.py:121-143."""
    lmm = lm.LinearMixedModel(phenotypes, ctx=ctx)
    lmm.add_random_effect(k)
    eig_L = lmm._get_eigen_L_()
    eig_R = lmm._get_eigen_R_(X=lmm.X)
    res = lmm.get_estimates(eig_L, method='REML', eig_R=eig_R)
    return lmm, eig_L, res


def run_emmax(hdf5_filename='/home/bv25/data/Ls154/Ls154_12.hdf5',
              out_file='/home/bv25/data/Ls154/Ls154_results.hdf5',
              min_maf=0.1, recalculate_kinship=True, chunk_size=1000, ctx=None):
    """
    Apply the EMMAX algorithm to hdf5 formated genotype/phenotype data (hdf5_data.py:70-187).
    """
    ctx = ctx or _lib.get_context()
    ih5f, in_opened = _open(hdf5_filename, 'r')
    gg = ih5f['genot_data']
    ig = ih5f['indiv_data']
    n_indivs = len(_arr(ig['indiv_ids']))

    if recalculate_kinship:
        k, n_snps = _ibd_kinship_device(ctx, gg, n_indivs, min_maf=min_maf)
    else:
        assert 'kinship' in ih5f.keys(), 'Kinship is missing.  Please calculate that first!'
        k = _arr(ih5f['kinship'])

    phenotypes = _arr(ig['phenotypes'])
    lmm, eig_L, res = _fit_null(ctx, k, phenotypes)

    oh5f, out_opened = _open(out_file, 'w')
    _put(oh5f, 'pseudo_heritability', np.array(res['pseudo_heritability']))
    _put(oh5f, 've', np.array(res['ve']))
    _put(oh5f, 'vg', np.array(res['vg']))
    _put(oh5f, 'max_ll', np.array(res['max_ll']))
    _put(oh5f, 'num_snps', _arr(ih5f['num_snps']))
    chrom_res_group = _group(oh5f, 'chrom_results')

    for chrom in gg.keys():
        crg = _group(chrom_res_group, chrom)
        freqs = _arr(gg[chrom]['freqs'])
        mafs = np.minimum(freqs, 1 - freqs)
        maf_filter = mafs > min_maf
        snps = _arr(gg[chrom]['raw_snps'])[maf_filter]
        positions = _arr(gg[chrom]['positions'])[maf_filter]
        r = lmm._emmax_f_test_(snps, res['H_sqrt_inv'], with_betas=False, emma_num=0, eig_L=eig_L)
        _put(crg, 'ps', r['ps'])
        _put(crg, 'positions', positions)
        if hasattr(oh5f, 'flush'):
            oh5f.flush()

    if in_opened:
        ih5f.close()
    if out_opened:
        oh5f.close()


def run_emmax_perm(hdf5_filename='/home/bv25/data/Ls154/Ls154_12.hdf5',
                   out_file='/home/bv25/data/Ls154/Ls154_results_perm.hdf5',
                   min_maf=0.1, recalculate_kinship=True, chunk_size=1000,
                   num_perm=500, ctx=None):
    """
    EMMAX plus the permutation-based genome-wide threshold (hdf5_data.py:191-351): scans every
    chromosome, then runs _emmax_permutations_ on all chromosomes but the last (:294,310-311,330).
    """
    ctx = ctx or _lib.get_context()
    ih5f, in_opened = _open(hdf5_filename, 'r')
    gg = ih5f['genot_data']
    ig = ih5f['indiv_data']
    n_indivs = len(_arr(ig['indiv_ids']))

    Kd, n_snps_k = _ibd_kinship_device(ctx, gg, n_indivs, min_maf=min_maf)
    k = Kd.download()

    oh5f, out_opened = _open(out_file, 'w')
    _put(oh5f, 'kinship', k)                                  # :241-243

    chromosomes = list(gg.keys())
    n_snps = 0
    for chrom in chromosomes:
        freqs = _arr(gg[chrom]['freqs'])
        mafs = np.minimum(freqs, 1 - freqs)
        n_snps = int(np.sum(mafs > min_maf))                  # :253 (the reference stores the LAST chromosome's count, :289)

    phenotypes = _arr(ig['phenotypes'])
    lmm, eig_L, res = _fit_null(ctx, Kd, phenotypes)

    _put(oh5f, 'pseudo_heritability', np.array(res['pseudo_heritability']))
    _put(oh5f, 've', np.array(res['ve']))
    _put(oh5f, 'vg', np.array(res['vg']))
    _put(oh5f, 'max_ll', np.array(res['max_ll']))
    _put(oh5f, 'num_snps', np.array(n_snps))
    chrom_res_group = _group(oh5f, 'chrom_results')

    chr12 = []
    for chrom in chromosomes:
        crg = _group(chrom_res_group, chrom)
        freqs = _arr(gg[chrom]['freqs'])
        mafs = np.minimum(freqs, 1 - freqs)
        maf_filter = mafs > min_maf
        snps = _arr(gg[chrom]['raw_snps'])[maf_filter]
        positions = _arr(gg[chrom]['positions'])[maf_filter]
        if chrom != chromosomes[-1]:
            chr12.append(snps)
        r = lmm._emmax_f_test_(snps, res['H_sqrt_inv'], with_betas=False, emma_num=0, eig_L=eig_L)
        _put(crg, 'ps', r['ps'])
        _put(crg, 'positions', positions)
        if hasattr(oh5f, 'flush'):
            oh5f.flush()

    chr12_snps = np.concatenate(chr12, axis=0) if chr12 else np.zeros((0, n_indivs), dtype=np.int8)
    perm_res = lmm._emmax_permutations_(chr12_snps, k, res['H_sqrt_inv'], num_perm=num_perm)     # :330

    perm_res['min_ps'].sort()                                 # :339
    perm_res['max_f_stats'].sort()                            # :340 (the reference's [::-1] on :341 is a no-op)
    five_perc_i = int(num_perm / 20)
    _put(oh5f, 'perm_min_ps', perm_res['min_ps'])
    _put(oh5f, 'perm_max_f_stats', perm_res['max_f_stats'])
    _put(oh5f, 'five_perc_perm_min_ps', perm_res['min_ps'][five_perc_i])
    _put(oh5f, 'five_perc_perm_max_f_stats', perm_res['max_f_stats'][five_perc_i])

    if in_opened:
        ih5f.close()
    if out_opened:
        oh5f.close()
