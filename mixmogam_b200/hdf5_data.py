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

__all__ = ['calculate_ibd_kinship', 'run_emmax', 'run_emmax_perm', 'stream_snps', 'write_genotype_file']


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


def stream_snps(ctx, parts, chunk_rows=None):
    """
    HDF5 dataset(s) -> resident genotype block without a host copy of the whole matrix (the reference reads
    `gg[chrom]['raw_snps'][...]` in one piece and then filters it, hdf5_data.py:162-169, :298-305).  `parts` is a list of
    (dataset, keep) pairs -- dataset: anything sliceable by rows that yields int8 [rows x n] (an h5py dataset decompresses its
    lzf chunks slice by slice; a numpy array works too), keep: boolean row filter or None.  A reader thread fills two
    page-locked chunks while the previous one is on its way over PCIe (mmg_snps_write).  Returns a ResidentSnps handle that
    the scans accept in place of `snps`.
    """
    import threading
    import queue
    parts = [(ds, None if keep is None else np.asarray(keep, dtype=bool)) for ds, keep in parts]
    for ds, keep in parts:
        if keep is not None and keep.shape[0] != int(ds.shape[0]):
            raise ValueError('row filter of %d entries for a dataset of %d rows' % (keep.shape[0], int(ds.shape[0])))
    n = int(parts[0][0].shape[1])
    total = sum(int(ds.shape[0]) if keep is None else int(keep.sum()) for ds, keep in parts)
    if total == 0:
        raise ValueError('no SNPs')
    if chunk_rows is None:
        chunk_rows = max(256, (64 << 20) // max(n, 1))
    ctx.snps_reserve(total, n)
    # the two page-locked chunk buffers are kept on the context between calls (allocating them costs ~50 ms each; run_emmax_perm
    # streams every chromosome twice)
    ring = getattr(ctx, '_stream_ring', None)
    if ring is None or ring[0].size < chunk_rows * n:
        ring = [_lib.pinned_empty((chunk_rows * n,), np.int8) for _ in range(2)]
        ctx._stream_ring = ring
    bufs = [b[:chunk_rows * n].reshape(chunk_rows, n) for b in ring]
    free_q, full_q = queue.Queue(), queue.Queue()
    for b in bufs:
        free_q.put(b)

    def reader():
        try:
            for ds, keep in parts:
                for r0 in range(0, int(ds.shape[0]), chunk_rows):
                    r1 = min(int(ds.shape[0]), r0 + chunk_rows)
                    if keep is not None and not keep[r0:r1].any():
                        continue
                    block = np.asarray(ds[r0:r1])
                    if keep is not None:
                        block = block[keep[r0:r1]]
                    if block.dtype != np.int8:
                        block = _lib._as_int8(block)
                    buf = free_q.get()
                    buf[:block.shape[0]] = block
                    full_q.put((buf, block.shape[0]))
            full_q.put(None)
        except BaseException as e:           # surfaces in the consumer
            full_q.put(e)

    th = threading.Thread(target=reader, daemon=True)
    th.start()
    row0 = 0
    while True:
        item = full_q.get()
        if item is None:
            break
        if isinstance(item, BaseException):
            raise item
        buf, k = item
        ctx.snps_write(row0, buf[:k])        # returns when the copy has left the buffer
        row0 += k
        free_q.put(buf)
    th.join()
    assert row0 == total
    return _lib.ResidentSnps(total, n)


def write_genotype_file(target, chromosomes, indiv_ids, phenotypes, sex=None, compression='lzf'):
    """
    Writer side of the layout hdf5_data reads (the reference's plink2hdf5.py:25-28, :57-59, :111-118, :216-226; the PLINK text
    parsing in front of it is out of scope):
        genot_data/chrom_<c>/{raw_snps int8 (m_c x n), positions, freqs, snp_ids, nts, nt_counts, missing_counts}
        indiv_data/{indiv_ids, sex, phenotypes},  num_snps
    `chromosomes`: mapping chrom -> dict with 'raw_snps' (int8 [m_c x n]) and optionally 'positions', 'snp_ids', 'nts',
    'nt_counts', 'missing_counts'; 'freqs' defaults to the allele frequency mean(raw_snps) / 2 (plink2hdf5.py:96-99).
    `target`: file name (h5py) or a mapping to fill.
    """
    f, opened = _open(target, 'w')

    def put(g, name, data):
        if hasattr(g, 'create_dataset'):
            g.create_dataset(name, data=data, **({'compression': compression} if compression and np.ndim(data) > 0 else {}))
        else:
            g[name] = np.asarray(data)

    gg = _group(f, 'genot_data')
    ig = _group(f, 'indiv_data')
    put(ig, 'indiv_ids', np.asarray(indiv_ids))
    put(ig, 'sex', np.asarray(sex if sex is not None else np.zeros(len(indiv_ids), dtype=np.int64)))
    put(ig, 'phenotypes', np.asarray(phenotypes, dtype=np.float64))
    tot = 0
    for chrom, d in chromosomes.items():
        raw = np.asarray(d['raw_snps'])
        if raw.dtype != np.int8:
            raw = _lib._as_int8(raw)
        m_c = raw.shape[0]
        name = str(chrom) if str(chrom).startswith('chrom_') else 'chrom_%s' % chrom
        cg = _group(gg, name)
        put(cg, 'raw_snps', raw)
        put(cg, 'positions', np.asarray(d.get('positions', np.arange(1, m_c + 1))))
        put(cg, 'freqs', np.asarray(d['freqs']) if 'freqs' in d else raw.mean(axis=1, dtype=np.float64) / 2.0)
        put(cg, 'snp_ids', np.asarray(d.get('snp_ids', np.array(['%s_%d' % (name, i) for i in range(m_c)], dtype='S'))))
        for opt in ('nts', 'nt_counts', 'missing_counts'):
            if opt in d:
                put(cg, opt, np.asarray(d[opt]))
        tot += m_c
    put(f, 'num_snps', np.array(tot))
    if opened:
        f.close()
    return f


def _ibd_kinship_device(ctx, gg, n_indivs, min_maf=None, use_normalized=False):
    """hdf5_data.py:30-62 / :84-115 / :205-237: K = sum over chromosomes of Z'Z / n_snps, then the inline
    scale_k.  Returns (K DeviceMatrix, n_snps)."""
    K = ctx.matrix(n_indivs, n_indivs)
    n_snps = 0
    for chrom in gg.keys():
        cg = gg[chrom]
        mask = None
        if min_maf is not None:
            freqs = _arr(cg['freqs'])
            mafs = np.minimum(freqs, 1 - freqs)
            mask = (mafs > min_maf)                          # :91-96
        m_c, n = stream_snps(ctx, [(cg['raw_snps'], None)]).shape
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
        positions = _arr(gg[chrom]['positions'])[maf_filter]
        snps = stream_snps(ctx, [(gg[chrom]['raw_snps'], maf_filter)])       # :162-169, chunk by chunk into HBM
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
        positions = _arr(gg[chrom]['positions'])[maf_filter]
        if chrom != chromosomes[-1]:
            chr12.append((gg[chrom]['raw_snps'], maf_filter))
        snps = stream_snps(ctx, [(gg[chrom]['raw_snps'], maf_filter)])
        r = lmm._emmax_f_test_(snps, res['H_sqrt_inv'], with_betas=False, emma_num=0, eig_L=eig_L)
        _put(crg, 'ps', r['ps'])
        _put(crg, 'positions', positions)
        if hasattr(oh5f, 'flush'):
            oh5f.flush()

    chr12_snps = stream_snps(ctx, chr12)                      # :294,:310-311 -- all chromosomes but the last, one resident block
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
