"""
Drop-in for mixmogam's `kinship` module, hot-path subset (reference kinship.py):

    calc_ibs_kinship(snps, snps_data_format='binary', snp_dtype='int8', dtype='single',
                     chunk_size=None, scaled=True)                          kinship.py:14-56
    calc_ibd_kinship(snps, dtype='single', scaled=True)                     kinship.py:59-75
    scale_k(k, verbose=False)                                               kinship.py:94-100
    prepare_k(k, k_accessions, accessions)                                  kinship.py:79-90
    update_k_monomorphic(...)                                               kinship.py:134-142
    load_kinship_from_file / save_kinship_to_file / save_kinship_in_text_format     kinship.py:145-174

Same names, argument meaning and return types; the arithmetic runs on the B200 through
libmixmogam_b200 (no CPU fallback):
  * IBS: the contraction over SNPs is an int8 tensor-core Gram (tcgen05.mma kind::i8, int32
    accumulators) of the 2x-1 coded ('binary') or thermometer coded ('diploid_int') genotypes; the
    unscaled kinship is bit-identical to the reference's float64 / float32-quotient result.
  * IBD: per-SNP standardisation kernel + FP64 accumulation (the reference accumulates in float32).
`dtype`, `snp_dtype` and `chunk_size` are accepted for signature compatibility; results are float64
(which is also what the reference returns: kinship.py:44,51 promote to float64).
"""
import numpy as np

from . import _lib

__all__ = ['calc_ibs_kinship', 'calc_ibd_kinship', 'scale_k', 'calc_ibs_kinship_device', 'partial_ibs_gram', 'prepare_k',
           'update_k_monomorphic', 'load_kinship_from_file', 'save_kinship_to_file', 'save_kinship_in_text_format']


def _coding(snps_data_format):
    if snps_data_format == 'binary':
        return _lib.CODING_BINARY
    if snps_data_format == 'diploid_int':
        return _lib.CODING_DIPLOID
    raise NotImplementedError          # kinship.py:45-46


def calc_ibs_kinship_device(snps, snps_data_format='binary', scaled=True, impl='auto', ctx=None):
    """As calc_ibs_kinship but leaves K on the device (returns a DeviceMatrix)."""
    ctx = ctx or _lib.get_context()
    coding = _coding(snps_data_format)
    m, n = ctx.kinship_gram_from(snps, coding, impl=impl)      # host genotypes stream in underneath the Gram
    K, _ = ctx.kinship_finalize(coding, m, scaled)
    return K


def calc_ibs_kinship(snps, snps_data_format='binary', snp_dtype='int8', dtype='single',
                     chunk_size=None, scaled=True, impl='auto', ctx=None):
    """
    Calculates IBS kinship (kinship.py:14-56).

    data_format: 'binary' (0/1 genotypes) and 'diploid_int' (0/1/2) are supported.
    Returns an np.matrix for 'binary' and an ndarray for 'diploid_int', float64, as the reference does.
    The array lives in page-locked host memory and is READ-ONLY (copy it to modify it): the library keeps the
    device copy it was downloaded from, so handing it to LinearMixedModel.add_random_effect / emmax costs no upload.
    """
    ctx = ctx or _lib.get_context()
    K = calc_ibs_kinship_device(snps, snps_data_format, scaled, impl, ctx)
    k_mat = K.download(pinned=True)         # page-locked: D2H at the PCIe rate
    ctx.remember_resident(k_mat, K)         # emmax(snps, y, K) right after this call finds K still in HBM
    if snps_data_format == 'binary':
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter('ignore', PendingDeprecationWarning)
            return np.asmatrix(k_mat)       # kinship.py:43-44: sm is a matrix, so is the result
    return k_mat


def partial_ibs_gram(snps, snps_data_format='binary', impl='auto', ctx=None, reset=True):
    """Multi-GPU building block: integer Gram of this rank's SNP slice, left resident for an int32
    all-reduce (mixmogam_b200.parallel.allreduce_gram).  Stream ordered: nothing waits for the Gram here."""
    ctx = ctx or _lib.get_context()
    if reset:
        ctx.kinship_gram_from(snps, _coding(snps_data_format), impl=impl)
    else:
        ctx.ensure_snps(snps)
        ctx.kinship_gram(_coding(snps_data_format), impl=impl, reset=False)


def calc_ibd_kinship(snps, dtype='single', scaled=True, ctx=None):
    """kinship.py:59-75.  A monomorphic SNP raises AssertionError like the reference's `assert` (:67)."""
    ctx = ctx or _lib.get_context()
    m, n = ctx.ensure_snps(snps)
    K = ctx.matrix(n, n)
    try:
        ctx.kinship_ibd_accumulate(K, 0, m)
    except _lib.MmgError as e:
        if e.code == -7:
            raise AssertionError('WTF?')    # kinship.py:67
        raise
    k_mat = K.download()
    K.free()
    k_mat = k_mat / float(m)                # :72
    if scaled:
        k_mat = scale_k(k_mat, ctx=ctx)
    return k_mat


def scale_k(k, verbose=False, ctx=None):
    """kinship.py:94-100: K * (n-1) / (tr K - sum(K)/n), evaluated on the device."""
    ctx = ctx or _lib.get_context()
    is_matrix = isinstance(k, np.matrix)
    if isinstance(k, _lib.LazyHostArray):
        k = k.host()
    K = _lib.DeviceMatrix.from_host(ctx, np.asarray(k, dtype=np.float64))
    scalar = ctx.scale_k(K)
    if verbose:
        print('Kinship scaled by: %0.4f' % scalar)
    out = K.download()
    K.free()
    if is_matrix:
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter('ignore', PendingDeprecationWarning)
            return np.asmatrix(out)
    return out


# ----------------------------------------------------------------------------------------------
# Kinship files (kinship.py:79-90, 134-174).  `kinship_file` is a file name (HDF5 through h5py, imported lazily: it is not part
# of this image) or any mapping with the same three datasets -- 'kinship' (n x n), 'accessions' (n), 'n_snps' (scalar).

def prepare_k(k, k_accessions, accessions):
    """kinship.py:79-90: the sub-matrix of k for `accessions`, in that order; accessions k does not know are skipped."""
    k_accessions = list(k_accessions)
    accessions = list(accessions)
    if k_accessions == accessions:
        return np.asmatrix(np.asarray(k))
    pos = {}
    for i, acc in enumerate(k_accessions):
        pos.setdefault(acc, i)                         # list.index semantics: the first occurrence
    indices_to_keep = [pos[acc] for acc in accessions if acc in pos]
    k = np.asarray(k)[indices_to_keep, :][:, indices_to_keep]
    return np.asmatrix(k)


def update_k_monomorphic(n_removed_snps, full_kinship, full_indivs, full_num_snps, retained_indivs, kinship_type='ibs', dtype='single'):
    """kinship.py:134-142: the IBS kinship after dropping SNPs that are monomorphic among the retained individuals."""
    assert kinship_type == 'ibs', 'Only IBS kinships can be updated at the moment'
    cut_kinship = prepare_k(full_kinship, full_indivs, retained_indivs)
    num_lines = cut_kinship.shape[0]
    m = np.ones((num_lines, num_lines), dtype=np.float32 if dtype == 'single' else np.float64) * n_removed_snps
    return (cut_kinship * full_num_snps - m) / (full_num_snps - n_removed_snps)


def _open_kinship(kinship_file, mode):
    if isinstance(kinship_file, str):
        try:
            import h5py
        except ImportError:
            raise ImportError('h5py is required to open %r; pass an in-memory mapping with the same datasets instead' % kinship_file)
        return h5py.File(kinship_file, mode), True
    return kinship_file, False


def load_kinship_from_file(kinship_file, accessions=None, scaled=True, ctx=None):
    """kinship.py:145-160.  Returns {'k', 'accessions', 'n_snps'}; scaling runs on the device (scale_k)."""
    if isinstance(kinship_file, str):
        import os
        assert os.path.isfile(kinship_file), 'File not found.'
    f, opened = _open_kinship(kinship_file, 'r')
    k = np.asarray(f['kinship'][...])
    k_accessions = list(np.asarray(f['accessions'][...]))
    n_snps = int(np.asarray(f['n_snps'][...]))
    if opened:
        f.close()
    if accessions:
        k = prepare_k(k, k_accessions, accessions)
    if scaled:
        k = scale_k(np.asarray(k, dtype=np.float64), ctx=ctx)
    return {'k': k, 'accessions': k_accessions, 'n_snps': n_snps}


def save_kinship_to_file(kinship_file, kinship_mat, k_accessions, n_snps):
    """kinship.py:164-169."""
    f, opened = _open_kinship(kinship_file, 'w')
    for name, data in (('kinship', np.asarray(kinship_mat)), ('accessions', np.asarray(k_accessions)), ('n_snps', np.asarray(n_snps))):
        if hasattr(f, 'create_dataset'):
            f.create_dataset(name, data=data)
        else:
            f[name] = data
    if opened:
        f.close()


def save_kinship_in_text_format(filename, k, accessions):
    """kinship.py:172-175: one line per accession, `acc,k_1,...,k_n`."""
    with open(filename, 'w') as f:
        for acc, row in zip(accessions, np.asarray(k)):
            f.write('%s,%s\n' % (acc, ','.join(map(str, row.tolist()))))
