"""GPU: stage 1 (kinship) through the C ABI against the oracle / golden vectors.  Bit-exact for the integer
Gram and the unscaled kinship (both codings, both Gram implementations)."""
import numpy as np
import pytest

from conftest import golden

pytestmark = pytest.mark.gpu


def _gram_ref(snps, coding):
    """Integer Gram of the coded planes through the FP64 BLAS: every sum is an integer far below 2^53, so the result is exact."""
    x = snps.astype(np.float64)
    if coding == 0:
        s = 2.0 * x - 1.0
        return np.rint(s.T @ s).astype(np.int64)
    t = np.concatenate([(x >= 1), (x >= 2)], axis=0).astype(np.float64)
    return np.rint(t.T @ t).astype(np.int64)


def _rand_snps(m, n, coding, seed):
    rng = np.random.default_rng(seed)
    if coding == 0:
        return (rng.random((m, n)) < rng.uniform(0.05, 0.95, size=(m, 1))).astype(np.int8)
    return rng.binomial(2, rng.uniform(0.05, 0.95, size=(m, 1)), size=(m, n)).astype(np.int8)


def _gram_kind(monkeypatch, impl):
    """'tcgen05' is the e2m1 / kind::mxf4 Gram as a CTA-pair MMA (default), 'tcgen05_mcast' its multicast form (MMG_GRAM_PAIR=0),
    'tcgen05_i8' the int8 one (MMG_GRAM_KIND=i8)."""
    if impl == 'tcgen05_i8':
        monkeypatch.setenv('MMG_GRAM_KIND', 'i8')
        return 'tcgen05'
    if impl == 'tcgen05_mcast':
        monkeypatch.setenv('MMG_GRAM_PAIR', '0')
        return 'tcgen05'
    return impl


@pytest.mark.parametrize('impl', ['simt', 'tcgen05', 'tcgen05_mcast', 'tcgen05_i8'])
@pytest.mark.parametrize('coding', [0, 1])
@pytest.mark.parametrize('m,n', [(700, 37), (3000, 198), (5001, 300), (1, 5), (129, 257), (2048, 1000), (65537, 260), (300, 512), (4097, 769)])
def test_gram_bit_exact(ctx, monkeypatch, impl, coding, m, n):
    kind = impl
    impl = _gram_kind(monkeypatch, impl)
    snps = _rand_snps(m, n, coding, seed=m * 1000 + n + coding)
    ctx.invalidate_snps()
    ctx.ensure_snps(snps)
    ctx.kinship_gram(coding, impl=impl)
    assert ctx.last_kernel_ms('gram_is_fp4') == (1.0 if kind in ('tcgen05', 'tcgen05_mcast') else 0.0)
    assert ctx.last_kernel_ms('gram_is_pair') == (1.0 if kind == 'tcgen05' else 0.0)
    G = ctx.kinship_gram_download()
    assert np.array_equal(G.astype(np.int64), _gram_ref(snps, coding))


@pytest.mark.parametrize('impl', ['simt', 'tcgen05', 'tcgen05_i8', 'tcgen05_overlap'])
def test_gram_chunk_boundary_and_accumulate(ctx, monkeypatch, impl):
    """m crosses the 65536-SNP pack chunk (three chunks with the packs of the later ones on the side stream under the previous
    Gram for 'tcgen05_overlap'); a second call with reset=False accumulates (multi-call == one call)."""
    n, m = 130, 70000
    if impl == 'tcgen05_overlap':
        monkeypatch.setenv('MMG_GRAM_OVERLAP', '1')
        impl, m = 'tcgen05', 140000
    impl = _gram_kind(monkeypatch, impl)
    snps = _rand_snps(m, n, 1, seed=5)
    ctx.invalidate_snps()
    ctx.ensure_snps(snps)
    ctx.kinship_gram(1, impl=impl)
    G1 = ctx.kinship_gram_download()
    assert np.array_equal(G1.astype(np.int64), _gram_ref(snps, 1))
    ctx.kinship_gram(1, impl=impl, snp_begin=0, snp_count=30001, reset=True)
    ctx.kinship_gram(1, impl=impl, snp_begin=30001, snp_count=m - 30001, reset=False)
    assert np.array_equal(ctx.kinship_gram_download(), G1)


@pytest.mark.parametrize('pinned', [False, True])
@pytest.mark.parametrize('coding,m,n', [(1, 140001, 130), (0, 65536, 77), (1, 300, 200), (0, 196608 + 5, 40)])
def test_gram_streamed_from_host_bit_exact(ctx, coding, m, n, pinned):
    """mmg_kinship_gram_i8_host: the genotypes stream into the resident block chunk by chunk while the Gram of the chunks
    that have landed runs; same integer Gram as upload-then-Gram, from pageable and from page-locked rows, across the
    65 536-SNP chunk boundaries (one chunk exactly, three chunks + 5, ragged), and the block is left resident."""
    from mixmogam_b200 import _lib
    snps = _rand_snps(m, n, coding, seed=9 * m + n)
    if pinned:
        host = _lib.pinned_empty((m, n), np.int8)
        host[...] = snps
        snps = host
    ctx.invalidate_snps()
    assert ctx.kinship_gram_from(snps, coding) == (m, n)
    packed, raw, _ = ctx.last_h2d_info()
    assert packed + raw == (m + 65535) // 65536 and (pinned or raw == 0)  # pageable rows all go through the 2-bit lane
    G = ctx.kinship_gram_download()
    ctx.kinship_gram(coding, reset=True)                                  # the resident copy, Gram again
    assert np.array_equal(ctx.kinship_gram_download(), G)
    if m <= 70000:
        assert np.array_equal(G.astype(np.int64), _gram_ref(np.asarray(snps), coding))
    else:
        x = np.asarray(snps)                                              # int64 reference in slabs (memory)
        ref = sum(_gram_ref(x[i:i + 50000], coding) for i in range(0, m, 50000))
        assert np.array_equal(G.astype(np.int64), ref)
    sums = ctx.snps_row_sums()
    assert np.array_equal(sums, np.asarray(snps).astype(np.int64).sum(axis=1))   # every row landed
    assert ctx.kinship_gram_from(snps, coding) == (m, n)                 # already resident: plain Gram, same result
    assert np.array_equal(ctx.kinship_gram_download(), G)


def test_gram_streamed_lanes(ctx, monkeypatch):
    """Both lanes of the streamed upload give the same resident block and Gram: packed lane off (MMG_H2D_PACK=0), and a
    chunk with a code the 2-bit packing cannot hold (5, invalid for the coding too) falling back to the raw lane and then
    being refused by the coding check."""
    from mixmogam_b200 import MmgError
    m, n = 150000, 70
    snps = _rand_snps(m, n, 1, seed=77)
    ctx.invalidate_snps()
    ctx.kinship_gram_from(snps, 1)
    assert ctx.last_h2d_info()[0] == 3
    G = ctx.kinship_gram_download()
    monkeypatch.setenv('MMG_H2D_PACK', '0')
    ctx.invalidate_snps()
    ctx.kinship_gram_from(snps, 1)
    assert ctx.last_h2d_info()[:2] == (0, 3)
    assert np.array_equal(ctx.kinship_gram_download(), G)
    monkeypatch.delenv('MMG_H2D_PACK')
    snps[70000, 3] = 5                                                    # second chunk
    ctx.invalidate_snps()
    with pytest.raises(MmgError):
        ctx.kinship_gram_from(snps, 1)
    assert ctx.last_h2d_info()[:2] == (1, 2)
    snps[70000, 3] = 0
    ctx.invalidate_snps()
    ctx.kinship_gram_from(snps, 1)
    snps2 = snps.copy()
    assert np.array_equal(ctx.snps_row_sums(), snps2.astype(np.int64).sum(axis=1))


def test_gram_streamed_rejects_out_of_domain_values(ctx):
    from mixmogam_b200 import MmgError
    snps = _rand_snps(70000, 40, 1, seed=2)
    snps[69999, 39] = 3
    ctx.invalidate_snps()
    with pytest.raises(MmgError):
        ctx.kinship_gram_from(snps, 1)
    snps[69999, 39] = 1
    ctx.kinship_gram_from(snps, 1)                                       # a failed call leaves nothing marked resident
    assert np.array_equal(ctx.kinship_gram_download().astype(np.int64), _gram_ref(snps, 1))


@pytest.mark.parametrize('coding', [0, 1])
def test_gram_split_k_tail_bit_exact(ctx, monkeypatch, coding):
    """n large enough that the tile count passes the number of co-resident clusters (n = 3112: 13 column tiles, 91 cluster
    tiles > 74): the tiles of the last partial wave are cut along K and accumulated with integer atomics.  Same bits as
    the unsplit schedule, the SIMT Gram and an exact FP64 BLAS reference (all sums < 2^53); two chunks, accumulate."""
    n, m = 3112, 66000
    snps = _rand_snps(m, n, coding, seed=31 + coding)
    ctx.invalidate_snps()
    ctx.ensure_snps(snps)
    ctx.kinship_gram(coding, impl='tcgen05')
    G = ctx.kinship_gram_download()
    monkeypatch.setenv('MMG_GRAM_SPLITK', '0')
    ctx.kinship_gram(coding, impl='tcgen05')
    assert np.array_equal(ctx.kinship_gram_download(), G)
    monkeypatch.delenv('MMG_GRAM_SPLITK')
    monkeypatch.setenv('MMG_GRAM_KIND', 'i8')                              # int8 operands: same integers as the e2m1 ones
    ctx.kinship_gram(coding, impl='tcgen05')
    assert ctx.last_kernel_ms('gram_is_fp4') == 0.0
    assert np.array_equal(ctx.kinship_gram_download(), G)
    monkeypatch.delenv('MMG_GRAM_KIND')
    x = snps.astype(np.float64)
    if coding == 0:
        s = 2.0 * x - 1.0
        ref = s.T @ s
    else:
        t1, t2 = (x >= 1).astype(np.float64), (x >= 2).astype(np.float64)
        ref = t1.T @ t1 + t2.T @ t2
    assert np.array_equal(G.astype(np.float64), ref)
    ctx.kinship_gram(coding, impl='tcgen05', snp_begin=0, snp_count=1000, reset=True)     # short K: too few K-blocks to split
    ctx.kinship_gram(coding, impl='tcgen05', snp_begin=1000, snp_count=m - 1000, reset=False)
    assert np.array_equal(ctx.kinship_gram_download(), G)


def test_gram_rejects_out_of_domain_values(ctx):
    from mixmogam_b200 import MmgError
    snps = _rand_snps(300, 40, 1, seed=1)
    snps[17, 3] = 3
    ctx.invalidate_snps()
    ctx.ensure_snps(snps)
    with pytest.raises(MmgError):
        ctx.kinship_gram(1)
    snps[17, 3] = 2
    ctx.invalidate_snps()
    ctx.ensure_snps(snps)
    with pytest.raises(MmgError):
        ctx.kinship_gram(0)          # a 2 is not a binary genotype


@pytest.mark.parametrize('impl', ['simt', 'tcgen05'])
@pytest.mark.parametrize('name,fmt', [('ibs_binary_n37.npz', 'binary'), ('ibs_diploid_n37.npz', 'diploid_int'),
                                      ('ibs_diploid_n198.npz', 'diploid_int')])
def test_calc_ibs_kinship_golden(ctx, impl, name, fmt):
    from mixmogam_b200 import kinship
    g = golden(name)
    ctx.invalidate_snps()
    K = kinship.calc_ibs_kinship(list(g['snps']), fmt, scaled=False, impl=impl)
    assert np.array_equal(np.asarray(K), g['K_unscaled'])            # bit-exact
    assert isinstance(K, np.matrix) == (fmt == 'binary')             # kinship.py:43 returns a matrix for 'binary'
    Ks = np.asarray(kinship.calc_ibs_kinship(g['snps'], fmt, scaled=True, impl=impl))
    ulp = np.abs(Ks - g['K_scaled']) / np.spacing(np.abs(g['K_scaled']))
    assert ulp.max() <= 4, ulp.max()
    with pytest.raises(NotImplementedError):
        kinship.calc_ibs_kinship(g['snps'], 'triploid')


def test_calc_ibs_kinship_matches_oracle_midsize(ctx):
    from mixmogam_b200 import kinship
    from oracle import reference_py3 as o
    snps = o.synth_genotypes(20000, 1000, 'diploid_int', seed=77)
    ctx.invalidate_snps()
    K = kinship.calc_ibs_kinship(snps, 'diploid_int', scaled=False)
    assert np.array_equal(K, o.calc_ibs_kinship_diploid_fast(snps, scaled=False))
    # size-independent properties: symmetry, unit diagonal, entries are multiples of 1/(2m) before the f32 quotient
    assert np.array_equal(K, K.T) and np.all(np.diag(K) == 1.0)


@pytest.mark.parametrize('ibd_impl,rtol,atol', [('tcgen05', 1e-8, 1e-10), ('dsyrk', 1e-11, 1e-13)])
def test_scale_k_and_ibd(ctx, monkeypatch, ibd_impl, rtol, atol):
    from mixmogam_b200 import kinship
    monkeypatch.setenv('MMG_IBD_IMPL', ibd_impl)
    g = golden('ibd_n37.npz')
    ctx.invalidate_snps()
    K = kinship.calc_ibd_kinship(list(g['snps']))
    np.testing.assert_allclose(K, g['K_double'], rtol=rtol, atol=atol)
    Ku = kinship.calc_ibd_kinship(g['snps'], scaled=False)
    np.testing.assert_allclose(Ku, g['K_double_unscaled'], rtol=rtol, atol=atol)
    np.testing.assert_allclose(kinship.scale_k(Ku), g['K_double'], rtol=rtol, atol=atol)
    # the float32-accumulating reference differs at the 1e-6 level only
    np.testing.assert_allclose(K, g['K_single'], rtol=2e-4, atol=2e-6)
    mono = g['snps'].copy()
    mono[3] = 1
    ctx.invalidate_snps()
    with pytest.raises(AssertionError):
        kinship.calc_ibd_kinship(mono)


@pytest.mark.parametrize('n,m', [(700, 9000), (300, 70000)])
def test_ibd_tensor_core_matches_fp64(ctx, monkeypatch, n, m):
    """int8 digit-plane IBD Gram (ibd_tc.cuh) against the FP64 library path and the float64 oracle: several row tiles,
    a ragged last K block, two packed chunks (m > 65536), rare alleles (wide range of 1/var weights), a MAF mask."""
    from mixmogam_b200 import kinship
    from oracle import reference_py3 as o
    rng = np.random.default_rng(n)
    f = np.concatenate([rng.uniform(0.004, 0.05, m // 4), rng.uniform(0.05, 0.5, m - m // 4)])[:, None]
    snps = ((rng.random((m, n)) < f).astype(np.int8) + (rng.random((m, n)) < f).astype(np.int8))
    snps = snps[snps.min(1) != snps.max(1)]
    ctx.invalidate_snps()
    monkeypatch.setenv('MMG_IBD_IMPL', 'tcgen05')
    Kt = kinship.calc_ibd_kinship(snps, scaled=False)
    assert ctx.last_kernel_ms('ibd') > 0                            # the tensor-core kernel ran
    monkeypatch.setenv('MMG_IBD_IMPL', 'dsyrk')
    Kd = kinship.calc_ibd_kinship(snps, scaled=False)
    scale = np.abs(Kd).max()
    assert np.max(np.abs(Kt - Kd)) <= 1e-9 * scale
    assert np.array_equal(Kt, Kt.T)
    if n * len(snps) <= 8e6:
        Ko = o.calc_ibd_kinship(list(snps), dtype='double', scaled=False)
        assert np.max(np.abs(Kt - Ko)) <= 1e-9 * scale
    # masked accumulation (hdf5_data.py:91-96)
    monkeypatch.setenv('MMG_IBD_IMPL', 'tcgen05')
    mask = (np.arange(len(snps)) % 3 != 0)
    Ka = ctx.matrix(n, n)
    used = ctx.kinship_ibd_accumulate(Ka, 0, len(snps), mask)
    assert used == int(mask.sum())
    monkeypatch.setenv('MMG_IBD_IMPL', 'dsyrk')
    Kb = ctx.matrix(n, n)
    ctx.kinship_ibd_accumulate(Kb, 0, len(snps), mask)
    a, b = Ka.download(), Kb.download()
    assert np.max(np.abs(a - b)) <= 1e-9 * np.abs(b).max()
    # genotypes outside {0,1,2} take the FP64 path transparently
    monkeypatch.setenv('MMG_IBD_IMPL', 'tcgen05')
    odd = snps[:500].copy()
    odd[odd == 2] = 3
    ctx.invalidate_snps()
    Ko3 = kinship.calc_ibd_kinship(odd, scaled=False)
    z = odd.astype(np.float64)
    z = (z - z.mean(1, keepdims=True)) / z.std(1, keepdims=True)
    np.testing.assert_allclose(Ko3, z.T @ z / len(odd), rtol=1e-10, atol=1e-12)


def test_returned_kinship_stays_resident(ctx):
    """calc_ibs_kinship hands back a read-only page-locked array and keeps the device copy: emmax(snps, y, K) must
    give the same result whether it finds that copy or uploads a fresh (writable) one."""
    from mixmogam_b200 import kinship, linear_models as lm
    from oracle import reference_py3 as o
    snps = o.synth_genotypes(3000, 300, 'diploid_int', seed=12)
    ctx.invalidate_snps()
    K = kinship.calc_ibs_kinship(snps, 'diploid_int')
    assert not K.flags.writeable and ctx.lookup_resident(K) is not None
    with pytest.raises(ValueError):
        K[0, 0] = 2.0
    y = o.synth_phenotype(snps, np.asarray(K), seed=4)
    r1 = lm.emmax(snps, y, K)
    K2 = np.array(K)                                     # a writable copy takes the upload path
    assert ctx.lookup_resident(K2) is None
    r2 = lm.emmax(snps, y, K2)
    assert np.array_equal(r1['ps'], r2['ps'])
    Kb = kinship.calc_ibs_kinship((snps > 0).astype(np.int8), 'binary')
    assert isinstance(Kb, np.matrix) and ctx.lookup_resident(Kb) is not None
    del K, Kb
