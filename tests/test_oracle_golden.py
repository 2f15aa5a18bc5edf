"""CPU: the oracle reproduces the committed golden vectors, and its own internal identities hold."""
import warnings

import numpy as np
import pytest

from conftest import golden
from oracle import reference_py3 as o

warnings.simplefilter('ignore')


def test_ibs_binary_literal():
    g = golden('ibs_binary_n37.npz')
    snps = g['snps']
    assert np.array_equal(o.calc_ibs_kinship(list(snps), 'binary', scaled=False), g['K_unscaled'])
    np.testing.assert_allclose(o.calc_ibs_kinship(list(snps), 'binary'), g['K_scaled'], rtol=1e-15)
    # 2-D array input == list input (hdf5_data.py:166-167 passes arrays)
    assert np.array_equal(o.calc_ibs_kinship(snps, 'binary', scaled=False), g['K_unscaled'])


def test_ibs_binary_integer_identity():
    """K = G_s/(2m) + 1/2 with the exact integer Gram G_s = S S', S = 2x-1 (what the device computes)."""
    g = golden('ibs_binary_n37.npz')
    s = 2 * g['snps'].astype(np.int64) - 1
    G = s.T @ s
    assert np.array_equal(G / (2.0 * len(s)) + 0.5, g['K_unscaled'])


def test_ibs_diploid_literal_and_identity():
    g = golden('ibs_diploid_n37.npz')
    snps = g['snps']
    lit = o.calc_ibs_kinship(list(snps), 'diploid_int', scaled=False)
    assert np.array_equal(lit, g['K_unscaled'])
    assert np.array_equal(o.calc_ibs_kinship_diploid_fast(snps, scaled=False), lit)
    assert np.all(np.diag(lit) == 1.0)                       # kinship.py:35 never fills the diagonal
    # thermometer Gram identity, as the device evaluates it
    t = np.concatenate([(snps >= 1), (snps >= 2)], axis=0).astype(np.int64)
    G = t.T @ t
    l1 = np.diag(G)[:, None] + np.diag(G)[None, :] - 2 * G
    c = len(snps) - 0.5 * l1
    k = (c.astype(np.float32) / np.float32(len(snps))).astype(np.float64)
    np.fill_diagonal(k, 1.0)
    assert np.array_equal(k, lit)


def test_ibs_chunking_is_irrelevant():
    g = golden('ibs_binary_n37.npz')
    a = o.calc_ibs_kinship(list(g['snps']), 'binary', scaled=False, chunk_size=64)
    assert np.array_equal(a, g['K_unscaled'])


def test_ibd():
    g = golden('ibd_n37.npz')
    np.testing.assert_allclose(o.calc_ibd_kinship(list(g['snps']), dtype='double'), g['K_double'], rtol=1e-13)
    np.testing.assert_allclose(o.calc_ibd_kinship(list(g['snps']), dtype='single'), g['K_single'], rtol=1e-6)
    np.testing.assert_allclose(g['K_single'], g['K_double'], rtol=2e-4, atol=2e-6)
    mono = g['snps'].copy()
    mono[3] = 1
    with pytest.raises(AssertionError):
        with np.errstate(all='ignore'):
            o.calc_ibd_kinship(list(mono))


def test_scale_k_idempotent():
    g = golden('ibs_binary_n37.npz')
    k = g['K_scaled']
    np.testing.assert_allclose(o.scale_k(k), k, rtol=1e-14)
    n = len(k)
    assert abs((np.trace(k) - k.sum() / n) / (n - 1) - 1.0) < 1e-13


def test_emmax_ft10_golden():
    g = golden('emmax_ft10_n198.npz')
    snps, y, K = g['snps'], g['y'], g['K']
    r = o.emmax(list(snps), y, K, dtype='double')
    np.testing.assert_allclose(r['ps'], g['double_ps'], rtol=1e-9)
    np.testing.assert_allclose(r['_delta'], g['double_delta'], rtol=1e-10)
    np.testing.assert_allclose(r['vg'], g['double_vg'], rtol=1e-9)
    # the float32-faithful mode stays within 1e-2 in -log10 p of the float64 algebra and ranks the top hits identically
    d = np.abs(np.log10(g['single_ps']) - np.log10(g['double_ps']))
    assert d.max() < 1e-2
    assert np.array_equal(np.argsort(g['single_ps'])[:20], np.argsort(g['double_ps'])[:20])


def test_closed_form_scan_matches_lstsq():
    """rss = h0_rss - (x~.y~)^2/(x~.x~), F = n_p r2/(1-r2): the algebra the fused kernel uses."""
    g = golden('emmax_ft10_n198.npz')
    snps, y, K = g['snps'][:500], g['y'], g['K']
    lmm = o.LinearMixedModel(y, 'double')
    lmm.add_random_effect(K)
    res = lmm.get_REML()
    H = res['H_sqrt_inv']
    r = lmm._emmax_f_test_(list(snps), H, emma_num=0)
    n = len(y)
    h0_X = H @ lmm.X
    Yt = H @ lmm.Y
    b, h0_rss = np.linalg.lstsq(h0_X, Yt, rcond=None)[:2]
    Yres = Yt - h0_X @ b
    Q = np.linalg.qr(h0_X)[0]
    R = (np.eye(n) - Q @ Q.T) @ H
    xt = snps.astype(np.float64) @ R.T
    xx = np.sum(xt * xt, axis=1)
    xy = xt @ Yres[:, 0]
    r2 = xy ** 2 / (xx * h0_rss[0])
    f = (n - 2) * r2 / (1 - r2)
    np.testing.assert_allclose(f, r['f_stats'], rtol=1e-8)
    # quadratic form on A = R'R (what the int8 path evaluates)
    A = R.T @ R
    xs = snps.astype(np.float64)
    np.testing.assert_allclose(np.einsum('si,ij,sj->s', xs, A, xs), xx, rtol=1e-10)


def test_vg_broadcast_quirk():
    g = golden('emmax_ft10_n198.npz')
    lmm = o.LinearMixedModel(g['y'], 'double')
    lmm.add_random_effect(g['K'])
    a = lmm.get_REML()
    eig_R = a['_eig_R']
    etas = eig_R['vectors'] @ lmm.Y
    p = len(eig_R['values'])
    closed = np.sum(etas ** 2) * np.sum(1.0 / (eig_R['values'] + a['delta'])) / p
    np.testing.assert_allclose(a['vg'], closed, rtol=1e-12)
    np.testing.assert_allclose(a['vg'], g['reml_vg'], rtol=1e-10)


def test_permutation_golden():
    g = golden('perm_n120.npz')
    lmm = o.LinearMixedModel(g['y'], 'double')
    lmm.add_random_effect(g['K'])
    res = lmm.get_REML()
    np.random.seed(int(g['seed']))
    pr = lmm._emmax_permutations_(g['snps'].astype(np.float64), g['K'], res['H_sqrt_inv'], num_perm=25)
    np.testing.assert_allclose(pr['_Ys'], g['Ys'], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(pr['min_ps'], g['min_ps'], rtol=1e-7)
