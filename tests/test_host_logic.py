"""CPU: the product's host/device headers (built with g++) against golden vectors; the C ABI loads and
exports every symbol of include/mixmogam_b200.h; the product never imports the oracle."""
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import ROOT, golden


def _run(built, mode, payload):
    exe = os.path.join(ROOT, 'tests', 'host_check')
    out = subprocess.run([exe, mode], input=np.asarray(payload, dtype=np.float64).tobytes(), capture_output=True, check=True)
    return np.frombuffer(out.stdout, dtype=np.float64)


@pytest.mark.parametrize('dfn,key', [(1, 'sf'), (2, 'sf2'), (3, 'sf3')])
def test_f_sf_matches_scipy(built, dfn, key):
    g = golden('f_sf.npz')
    F, D, S = g['f'].ravel(), g['dfd'].ravel(), g[key].ravel()
    payload = np.concatenate([[F.size], np.stack([F, np.full_like(F, dfn), D], axis=1).ravel()])
    out = _run(built, 'fsf', payload)
    tail = (S > 0) & (S < 0.5)
    assert np.max(np.abs(out[tail] - S[tail]) / S[tail]) < 1e-10       # relative in the tail, down to 1e-300
    body = S >= 0.5
    assert np.max(np.abs(out[body] - S[body])) < 1e-13
    assert np.all(out[S == 0] < 1e-300)                                # underflows where scipy returns 0
    # -log10 space, the parity measure
    assert np.max(np.abs(np.log10(out[tail]) - np.log10(S[tail]))) < 1e-9


def _reml_case(built, gname, ngrids):
    g = golden(gname)
    from oracle import reference_py3 as o
    lmm = o.LinearMixedModel(g['y'], 'double')
    lmm.add_random_effect(g['K'])
    res = lmm.get_REML(ngrids=ngrids)
    eig_R = res['_eig_R']
    etas = (eig_R['vectors'] @ lmm.Y).reshape(-1)
    deltas = np.asarray(res['_deltas'], dtype=np.float64)
    p = len(etas)
    payload = np.concatenate([[p, len(deltas), 1e-6], eig_R['values'], etas ** 2, deltas])
    out = _run(built, 'reml', payload)
    gl = len(deltas)
    np.testing.assert_allclose(out[3:3 + gl], res['_lls'], rtol=1e-10, atol=1e-9)
    np.testing.assert_allclose(out[3 + gl:3 + 2 * gl], res['_dlls'], rtol=1e-8, atol=1e-8)
    return out[0], out[1], int(out[2]), res


@pytest.mark.parametrize('ngrids', [100, 50, 10])
def test_reml_logic_interior_root(built, ngrids):
    """Sign change + secant refinement (linear_models.py:829-871) replicated step for step."""
    delta, ll, flags, res = _reml_case(built, 'emmax_diploid_n400.npz', ngrids)
    assert flags == 7
    assert abs(delta - res['delta']) / res['delta'] < 1e-9
    assert abs(ll - res['max_ll']) < 1e-8


def test_reml_logic_boundary_optimum(built):
    """FT10 against unrelated synthetic genotypes has no heritable signal: no sign change, the optimum is
    the last grid point delta = e^10 (linear_models.py:888-891)."""
    delta, ll, flags, res = _reml_case(built, 'emmax_ft10_n198.npz', 100)
    assert flags == 0
    assert delta == res['delta'] == np.exp(10.0)
    assert abs(ll - res['max_ll']) < 1e-8


def test_reml_logic_boundary_cases(built):
    """No sign change -> grid maximum (linear_models.py:888-891)."""
    p = 50
    rng = np.random.default_rng(0)
    eig = np.sort(rng.uniform(0.1, 3, p))
    deltas = np.exp(np.linspace(-10, 10, 51))
    sq = rng.normal(size=p) ** 2 * 1e-6 * (eig + 1e-9)      # essentially no genetic signal: ll increases with delta
    out = _run(built, 'reml', np.concatenate([[p, 51, 1e-6], eig, sq, deltas]))
    lls = out[3:3 + 51]
    dlls = out[3 + 51:]
    has_interval = np.any((dlls[1:] < 0) & (dlls[:-1] > 0))
    if not has_interval:
        assert int(out[2]) == 0
        assert out[0] == deltas[np.argmax(lls)]


def _declared(header):
    hdr = open(os.path.join(ROOT, 'include', header)).read()
    hdr = re.sub(r'/\*.*?\*/', '', hdr, flags=re.S)
    return set(re.findall(r'\b(mmg_[a-z0-9_]+)\s*\(', hdr))


def test_abi_exports_every_declared_symbol(built):
    declared = _declared('mixmogam_b200.h')
    assert len(declared) >= 40
    from mixmogam_b200 import _lib
    lib = _lib.load_library()
    for name in sorted(declared):
        assert hasattr(lib, name), 'library does not export %s' % name
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)


def test_bench_library_is_separate(built):
    """The microbenchmark kernels live in their own shared library; the product library exports none of them."""
    declared = _declared('mixmogam_b200_bench.h')
    from mixmogam_b200 import _lib
    assert declared == set(_lib.BENCH_SIGNATURES) and declared
    lib, blib = _lib.load_library(), _lib.load_bench_library()
    for name in sorted(declared):
        assert hasattr(blib, name), 'bench library does not export %s' % name
        assert not hasattr(lib, name), 'product library still exports %s' % name


def test_no_gpu_fails_loudly(built):
    """Without a device the product raises; it never falls back to CPU code."""
    import ctypes
    from mixmogam_b200 import _lib
    lib = _lib.load_library()
    h = ctypes.c_void_p(0)
    rc = lib.mmg_create(0, ctypes.byref(h))
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if not has_gpu:
        assert rc != 0 and not h.value
        assert b'no CPU fallback' in lib.mmg_last_error(None) or rc == -2
        with pytest.raises(_lib.MmgError):
            _lib.Context(0)
    else:
        assert rc == 0
        lib.mmg_destroy(h)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, 'mixmogam_b200')
    for dp, dn, fn in os.walk(pkg):
        for f in fn:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                src = open(os.path.join(dp, f)).read()
                assert 'oracle' not in src.replace('no oracle', ''), '%s references the oracle' % f
                assert 'import scipy' not in src and 'from scipy' not in src, '%s routes through scipy' % f


def test_base256_digit_expansion(built):
    """digits.cuh (the digit planes of the int8 scan): every digit fits an int8, no carry leaves the top digit, the
    expansion equals rint(r 256^S) EXACTLY (integer arithmetic), and every prefix of S' planes is within
    (128/255) 256^-S' of r -- the truncation bound the scan certifies -- including the edges of the accepted
    interval and the sliver (0.498, 1/2) that needs one more exponent bit."""
    from fractions import Fraction
    rng = np.random.default_rng(0)
    edge = [127.0 / 255.0, -127.0 / 255.0, np.nextafter(0.5, 0), -np.nextafter(0.5, 0), 0.25, -0.25, 0.499, 0.4981, 0.498,
            np.nextafter(0.498, 1), -0.498, 1.0, 3.0, -7.5, 1e-300, -1e300, 0.0, 2.0 ** -40, 127.5 / 256, -128.5 / 256, 255.0 / 512]
    vals = np.concatenate([edge, rng.standard_normal(3000) * 10.0 ** rng.uniform(-6, 6, 3000),
                           0.498 * 2.0 ** rng.integers(-5, 5, 500) * (1 - rng.uniform(0, 1e-6, 500))])
    bound = Fraction(128, 255)
    for S in (1, 4, 6, 7):
        out = _run(built, 'digits', np.concatenate([[vals.size, S], vals])).reshape(vals.size, S + 2)
        for a, row in zip(vals, out):
            E, digs, carry = int(row[0]), [int(d) for d in row[1:1 + S]], row[1 + S]
            assert carry == 0 and min(digs) >= -128 and max(digs) <= 127
            r = Fraction(float(a)) / Fraction(2) ** E
            if a != 0.0:
                assert Fraction(249, 1000) <= abs(r) <= Fraction(498, 1000)            # in range, no wasted head-room
            N = sum(d * 256 ** (S - 1 - k) for k, d in enumerate(digs))
            assert abs(r * 256 ** S - N) <= Fraction(1, 2)                              # N = rint(r 256^S)
            for Sp in range(0, S + 1):
                part = sum(Fraction(d, 256 ** (k + 1)) for k, d in enumerate(digs[:Sp]))
                assert abs(r - part) <= bound / 256 ** Sp


@pytest.mark.parametrize('rows,n,ld,threads', [(1, 1, 1, 1), (7, 37, 37, 3), (64, 10000, 10000, 4), (33, 198, 256, 2), (5, 31, 40, 8), (300, 131, 131, 5)])
def test_host_pack2_matches_numpy(built, rows, n, ld, threads):
    """The host side of the packed genotype upload (csrc/host_pack.cpp, AVX2 or scalar): 2 bits per code, code j of a row in
    bits 2 (j % 4) of byte j / 4, zeroed tail; any code outside 0..3 is reported."""
    from mixmogam_b200 import _lib
    lib = _lib.load_library()
    rng = np.random.default_rng(rows * 1000 + n)
    buf = rng.integers(0, 4, size=(rows, ld), dtype=np.int8)
    n4 = (n + 3) // 4
    dst_ld = (n4 + 15) // 16 * 16
    out = np.full((rows, dst_ld), 0xAB, dtype=np.uint8)
    rc = lib.mmg_host_pack2(buf.ctypes.data, rows, n, ld, out.ctypes.data, dst_ld, threads)
    assert rc == 0
    x = np.zeros((rows, n4 * 4), dtype=np.uint8)
    x[:, :n] = buf[:, :n]
    ref = x[:, 0::4] | (x[:, 1::4] << 2) | (x[:, 2::4] << 4) | (x[:, 3::4] << 6)
    assert np.array_equal(out[:, :n4], ref) and not out[:, n4:].any()
    for bad in (4, -1, 127, -128):
        b2 = buf.copy()
        b2[rows - 1, n - 1] = bad
        assert lib.mmg_host_pack2(b2.ctypes.data, rows, n, ld, out.ctypes.data, dst_ld, threads) == 1
    if ld > n:                                         # bytes between n and ld are not part of the row
        b3 = buf.copy()
        b3[:, n:] = 9
        assert lib.mmg_host_pack2(b3.ctypes.data, rows, n, ld, out.ctypes.data, dst_ld, threads) == 0
        assert np.array_equal(out[:, :n4], ref)
    assert lib.mmg_host_threads_default() >= 1


def test_int8_quad_form_error_bound_is_sound(built):
    """The int8 digit-plane product A = R'R of the scan (scan_tc.cuh: 7 base-256 planes of R 2^-F, plane pairs p + q < 7) in
    exact rational arithmetic, with the product's own digit expansion (digits.cuh through the host check program): the
    entry-wise error stays below ozaki_error_bound (mirrored here) -- the term the scan adds to its certified bound."""
    from fractions import Fraction
    P = L = 7
    rng = np.random.default_rng(3)
    n_out, n = 48, 6
    R = rng.standard_normal((n_out, n)) * 10.0 ** rng.uniform(-4, 0, size=(n_out, 1))
    R[0, 0] = np.abs(R).max() * 1.7                        # the entry that sets the global scale
    rmax = float(np.abs(R).max())
    out = _run(built, 'digits', np.concatenate([[1, 1], [rmax]]))
    F = int(out[0])                                        # digit256_exponent(rmax)
    scaled = np.ldexp(R, -F).ravel()
    assert np.abs(scaled).max() <= 0.498
    # digits of every entry with the library's own expansion: pass r 2^-F as "a" of an entry whose exponent comes out 0
    # (|r| in [0.249, 0.498]) is not general, so expand here with the same integer rule and cross-check a few through the binary
    def split(r):
        N = int(np.rint(np.ldexp(r, 8 * P)))
        digs = []
        for _ in range(P):
            d = ((N + 128) & 255) - 128
            digs.append(d)
            N = (N - d) >> 8
        assert N == 0
        return digs[::-1]
    big = [v for v in scaled if 0.249 <= abs(v) <= 0.498][:5]
    if big:
        chk = _run(built, 'digits', np.concatenate([[len(big), P], big])).reshape(len(big), P + 2)
        for v, row in zip(big, chk):
            assert int(row[0]) == 0 and [int(d) for d in row[1:1 + P]] == split(v)
    D = np.array([split(v) for v in scaled], dtype=object).reshape(n_out, n, P)
    rho = Fraction(128, 255) / Fraction(256) ** P
    dropped = sum(Fraction(min(s + 1, 2 * P - 1 - s) * 16384, 256 ** (s + 2)) for s in range(L, 2 * P - 1))
    bound = n_out * (dropped + rho + rho * rho)            # ozaki_error_bound(n_out), units of 2^2F
    worst = Fraction(0)
    for i in range(n):
        for j in range(i + 1):
            exact = sum(Fraction(float(scaled[k * n + i])) * Fraction(float(scaled[k * n + j])) for k in range(n_out))
            approx = Fraction(0)
            for p in range(P):
                for q in range(P):
                    if p + q < L:
                        g = sum(int(D[k, i, p]) * int(D[k, j, q]) for k in range(n_out))
                        approx += Fraction(g, 256 ** (p + q + 2))
            worst = max(worst, abs(exact - approx))
    assert worst <= bound
    assert worst > bound / 10 ** 6                          # and the bound is not vacuous


def test_scan_truncation_bound_is_sound(built):
    """The quadratic form of the int8 scan in exact rational arithmetic: x'Ax = sum_j A_jj x_j^2 + sum_j x_j sum_{i<j} 2 A_ji x_i
    with B = 2A 2^-E (strictly lower triangle) cut into 6 base-256 digit planes (digits.cuh rule) and only the first S used:
    the error never exceeds the bound the kernel certifies, (64/255) 256^-S 2^E ||x||_1^2 (scan_tc.cuh, api.cu set_bscale)."""
    from fractions import Fraction
    rng = np.random.default_rng(11)
    n, planes = 24, 6
    Rm = rng.standard_normal((n, n)) / np.sqrt(n)
    A = Rm.T @ Rm
    amax = max(abs(2 * A[j, i]) for j in range(n) for i in range(j))
    E = int(_run(built, 'digits', np.array([1, 1, amax]))[0])          # digit256_exponent(amax)

    def split(r):
        N = int(np.rint(np.ldexp(r, 8 * planes)))
        digs = []
        for _ in range(planes):
            d = ((N + 128) & 255) - 128
            digs.append(d)
            N = (N - d) >> 8
        assert N == 0
        return digs[::-1]

    dig = {(j, i): split(float(np.ldexp(2 * A[j, i], -E))) for j in range(n) for i in range(j)}
    two_E = Fraction(2) ** E
    for trial in range(6):
        x = rng.integers(0, 3, size=n) if trial else np.full(n, 2)
        l1 = int(np.abs(x).sum())
        exact = sum(Fraction(float(A[j, j])) * int(x[j]) ** 2 for j in range(n)) + \
            sum(Fraction(float(2 * A[j, i])) * int(x[j]) * int(x[i]) for j in range(n) for i in range(j))
        for S in range(1, planes + 1):
            q = Fraction(0)
            for k in range(S):
                acc = sum(int(x[j]) * sum(dig[(j, i)][k] * int(x[i]) for i in range(j)) for j in range(n))   # what the MMA + epilogue sum
                q += Fraction(acc, 256 ** (k + 1))
            approx = sum(Fraction(float(A[j, j])) * int(x[j]) ** 2 for j in range(n)) + q * two_E
            bound = Fraction(64, 255) / Fraction(256) ** S * two_E * l1 * l1
            assert abs(approx - exact) <= bound, (trial, S)


def test_packed_genotypes_roundtrip_host_only():
    """mmg_host_pack2 / PackedGenotypes: 2 bits per genotype, code j in bits 2 (j % 4) of byte j // 4; no GPU involved."""
    import mixmogam_b200 as mb
    rng = np.random.default_rng(3)
    for n in (1, 3, 4, 37, 198, 1001):
        x = rng.integers(0, 4, size=(57, n)).astype(np.int8)
        pk = mb.pack_genotypes(x)
        assert pk.shape == (57, n) and len(pk) == 57 and pk.packed.dtype == np.uint8 and not pk.packed.flags.writeable
        assert np.array_equal(pk.unpack(), x)
        assert np.array_equal(pk[5], x[5]) and np.array_equal(pk[[1, 7, 9]], x[[1, 7, 9]]) and np.array_equal(pk[10:20].unpack(), x[10:20])
        # the bit layout itself, spelled out
        j = n - 1
        assert (int(pk.packed[56, j // 4]) >> (2 * (j % 4))) & 3 == int(x[56, j])
    with pytest.raises(ValueError):
        mb.pack_genotypes(np.full((2, 8), 4, dtype=np.int8))


def test_kinship_file_helpers_mapping():
    """prepare_k / save / load (kinship.py:79-90, 145-169) on an in-memory mapping (h5py is not part of this image)."""
    from mixmogam_b200 import kinship
    rng = np.random.default_rng(0)
    a = rng.random((6, 6))
    k = a @ a.T
    acc = ['a%d' % i for i in range(6)]
    store = {}
    kinship.save_kinship_to_file(store, k, acc, 1234)
    d = kinship.load_kinship_from_file(store, scaled=False)
    assert np.array_equal(np.asarray(d['k']), k) and list(d['accessions']) == acc and d['n_snps'] == 1234
    sub = kinship.load_kinship_from_file(store, accessions=['a4', 'zz', 'a1'], scaled=False)['k']
    assert isinstance(sub, np.matrix) and np.array_equal(np.asarray(sub), k[[4, 1]][:, [4, 1]])
    assert np.array_equal(np.asarray(kinship.prepare_k(k, acc, acc)), k)
    upd = kinship.update_k_monomorphic(10, k, acc, 100, ['a0', 'a2'], dtype='double')
    np.testing.assert_allclose(np.asarray(upd), (k[[0, 2]][:, [0, 2]] * 100 - 10) / 90.0)


def test_write_genotype_file_layout():
    """The plink2hdf5 layout hdf5_data reads (plink2hdf5.py:25-28,57-59,111-118,226)."""
    from mixmogam_b200 import hdf5_data
    rng = np.random.default_rng(1)
    chroms = {1: {'raw_snps': rng.integers(0, 3, (20, 9)).astype(np.int8)}, 'chrom_2': {'raw_snps': rng.integers(0, 3, (5, 9)).astype(np.int8),
                                                                                         'positions': np.arange(5) * 7}}
    f = hdf5_data.write_genotype_file({}, chroms, np.arange(9), rng.standard_normal(9))
    assert set(f.keys()) == {'genot_data', 'indiv_data', 'num_snps'} and int(f['num_snps']) == 25
    assert set(f['genot_data'].keys()) == {'chrom_1', 'chrom_2'}
    c1 = f['genot_data']['chrom_1']
    assert {'raw_snps', 'positions', 'freqs', 'snp_ids'} <= set(c1.keys()) and c1['raw_snps'].dtype == np.int8
    np.testing.assert_allclose(c1['freqs'], chroms[1]['raw_snps'].mean(1) / 2.0)
    assert np.array_equal(f['genot_data']['chrom_2']['positions'], np.arange(5) * 7)
    assert set(f['indiv_data'].keys()) == {'indiv_ids', 'sex', 'phenotypes'}


def test_f_sf_series_branch_dense(built):
    """The power-series branch of f_sf (fdist.cuh: small F at large dfd -- where nearly every SNP of a scan lands) on a dense grid
    against scipy, including the switch-over to the continued fraction at (a + b) z = 4."""
    from scipy import stats
    for dfn in (1.0, 2.0):
        for dfd in (196.0, 397.0, 4998.0, 9998.0, 49998.0):
            F = np.concatenate([np.logspace(-14, 0, 150), np.linspace(1.0, 12.0, 221)])
            payload = np.concatenate([[F.size], np.stack([F, np.full_like(F, dfn), np.full_like(F, dfd)], axis=1).ravel()])
            out = _run(built, 'fsf', payload)
            ref = stats.f.sf(F, dfn, dfd)
            assert np.max(np.abs(out - ref) / ref) < 1e-11, (dfn, dfd)


def test_permutation_shuffle_of_the_column_view_matches_the_reference_call():
    """_emmax_permutations_ shuffles the 1-D view of the (n, 1) phenotype column: same draws from the legacy global RNG and the same
    cumulative permutations as the reference's `sp.random.shuffle(Y)` on the 2-D array (linear_models.py:1151-1154)."""
    n = 777
    a = np.random.RandomState(7).standard_normal((n, 1))
    b = a.copy()
    np.random.seed(20240601)
    ref = []
    for _ in range(4):
        np.random.shuffle(a)
        ref.append(a[:, 0].copy())
    tail_ref = np.random.random()
    np.random.seed(20240601)
    col = b[:, 0]
    for k in range(4):
        np.random.shuffle(col)
        assert np.array_equal(col, ref[k])
    assert np.random.random() == tail_ref                   # the RNG stream is in the same state afterwards
