"""
Drop-in for the EMMAX path of mixmogam's `linear_models` module (reference linear_models.py):

    LinearMixedModel(Y, dtype='single')                                      :558
        .add_factor / .set_factors                                           :98-130
        .add_random_effect(cov_matrix)                                       :577
        ._get_eigen_L_(K=None) / ._get_eigen_R_(X=None, K=None)              :589 / :600
        ._rell_ / ._redll_                                                   :618 / :626
        .get_REML(ngrids=100, llim=-10, ulim=10, esp=1e-6, eig_L, eig_R)     :653
        .get_estimates(eig_L, K, xs, ngrids=50, ..., method='REML', eig_R)   :771
        .expedited_REML_t_test(snps, ...)                                    :931
        .emmax_f_test(snps, snp_priors, Z, with_betas, method, eig_L, eig_R, emma_num=100)   :1233
        ._emmax_f_test_(snps, H_sqrt_inv, ...)                               :1272
        ._emmax_permutations_(snps, K, H_sqrt_inv, num_perm=100)             :1125
        .emmax_GxT_f_test(snps, E, Z, ...) / ._emmax_GxT_f_test_             :1383 / :1422
    emmax_w_two_env(snps, phenotypes, K, E, cofactors=None, Z=None)          :1749
    emmax(snps, phenotypes, K, cofactors=None, Z=None, with_betas=False, emma_num=0)          :1790
    emma(snps, phenotypes, K, cofactors=None)                                :1725
    emmax_multi(snps, phenotypes[T], K, cofactors=None)                      T x emmax() on one eigenbasis, one scan launch
    get_emma_reml_estimates(y, K, ...)                                       :1690 (single-K form)

Same names, arguments, defaults, returned dict keys and error behaviour.  The n x n objects (K, the
eigenbases, H_sqrt_inv, M) live in HBM; the three stages run in hand-written sm_100a kernels behind
libmixmogam_b200 (no CPU fallback).  All device arithmetic is float64, i.e. the reference's algebra at
higher precision than its hard-coded float32 (`dtype='single'`, :558,:589,:600,:773,:1283);
the `dtype` arguments are accepted and ignored.

Known, deliberate differences from the reference (each covered by a test):
  * eig_R is computed once per model, not recomputed inside get_estimates (:787 always recomputes it;
    the values are identical).
  * `vg`/`ve` reproduce the reference's (p,1)/(p,) broadcast at :894-896 in closed form:
    vg = sum(sq_etas) * sum(1/(lambda+delta)) / p.
  * eigenvectors are defined up to sign (cuSOLVER vs LAPACK): compare H'H, not H.
"""
import time
import warnings

import os

import numpy as np

from . import _lib
from . import kinship
from ._lib import DeviceMatrix, LazyHostArray, LazyScaledRows

__all__ = ['LinearModel', 'LinearMixedModel', 'emmax', 'emmax_multi', 'emmax_w_two_env', 'emma', 'get_emma_reml_estimates']

_VERBOSE = False


def _say(*a):
    if _VERBOSE:
        print(*a)


def _is_none(x):
    return x is None


def _col(x, n):
    a = np.asarray(x, dtype=np.float64)
    return a.reshape(n, -1) if a.ndim != 2 or a.shape[0] != n else a


class _LazyIdentity(object):
    """random_effects[0][1] of the reference is a dense identity (linear_models.py:574); keep it lazy."""

    def __init__(self, n):
        self.n = n
        self.shape = (n, n)

    def __array__(self, dtype=None, copy=None):
        return np.identity(self.n, dtype=dtype or np.float64)


class LinearModel(object):
    """linear_models.py:81-130 (the parts LinearMixedModel inherits on the EMMAX path)."""

    def __init__(self, Y=None):
        self.n = len(Y)
        self.Y = np.asarray(Y, dtype=np.float64).reshape(self.n, 1)
        self.X = np.ones((self.n, 1))
        self.p = 1
        self.beta_est = None
        self.cofactors = []

    def add_factor(self, x, lin_depend_thres=1e-8):
        """linear_models.py:98-113."""
        new_x = np.array(x, dtype=np.float64)
        new_x.shape = len(x)
        (beta, rss, rank, sigma) = np.linalg.lstsq(self.X, new_x, rcond=None)
        rss_v = float(rss[0]) if np.size(rss) else float(np.sum((new_x - self.X @ beta) ** 2))
        if rss_v < lin_depend_thres:
            warnings.warn('A factor was found to be linearly dependent on the factors already in the X matrix.  Hence skipping it!')
            return False
        new_x.shape = (self.n, 1)
        self.X = np.hstack([self.X, new_x])
        self.cofactors.append(x)
        self.p += 1
        self._invalidate()
        return True

    def set_factors(self, factors, include_intercept=True):
        """linear_models.py:116-130."""
        self.p = 0
        if include_intercept:
            self.X = np.ones((self.n, 1))
            self.p = 1
            if len(factors) > 0:
                self.X = np.hstack([self.X, np.asarray(factors, dtype=np.float64).T])
            self.p += len(factors)
        else:
            self.X = np.asarray(factors, dtype=np.float64).T
            self.p = len(factors)
        self._invalidate()

    def _invalidate(self):
        pass

    def fast_f_test(self, snps, verbose=True, Z=None, with_betas=False, ctx=None, scan_impl='auto'):
        """
        LM implementation, single SNPs (linear_models.py:196-257): the EMMAX scan without a kinship -- the rotation is
        M = I - QQ' alone (:215-218), the per-SNP least squares, F and p are the same fused kernel
        (_emmax_f_test_ with H = I).  Z is accepted and ignored, as in the reference.  Returns the reference's dict:
        ps, f_stats, rss, var_perc, h0_rss, h0_betas (+ betas with with_betas=True).
        """
        ctx = ctx or getattr(self, 'ctx', None) or _lib.get_context()
        mdl = LinearMixedModel(self.Y.reshape(-1), ctx=ctx, scan_impl=getattr(self, 'scan_impl', scan_impl))
        mdl.X = self.X
        mdl.p = self.p
        eye = ctx.matrix(self.n, self.n)            # zero-filled on the device
        ctx.add_diag(eye, 1.0)
        try:
            return mdl._emmax_f_test_(snps, eye, verbose=verbose, with_betas=with_betas, emma_num=0)
        finally:
            eye.free()


class EigenDict(dict):
    """{'values', 'vectors'} as the reference returns (linear_models.py:596,615), eigenvectors as ROWS.
    'vectors' is a LazyHostArray: it stays in HBM until someone asks for the numbers."""
    pass


class LinearMixedModel(LinearModel):
    """
    A class for linear mixed models (linear_models.py:554).
    """

    def __init__(self, Y=None, dtype='single', ctx=None, scan_impl='auto', shard=None):
        self.ctx = ctx or _lib.get_context()
        self.scan_impl = scan_impl
        self.shard = None
        if shard is not None:
            self.set_sharding(**(shard if isinstance(shard, dict) else {'group': shard}))
        self.n = len(Y)
        self.y_var = np.var(Y, ddof=1)
        self.Y = np.array(Y, dtype=np.float64).reshape(self.n, 1)
        self.X = np.ones((self.n, 1))
        self.p = 1
        self.beta_est = None
        self.cofactors = []
        # A list of random effect type, and the cov matrix.  The first random effect is the IID error.
        self.random_effects = [('normal', _LazyIdentity(self.n))]
        self._K_dev = []
        self._eig_R_cache = None

    def _invalidate(self):
        self._eig_R_cache = None

    def set_sharding(self, group='world', m_total=None):
        """Opt in to the SNP-sharded scan (one process per GPU, torch.distributed NCCL group): EVERY rank of `group` holds the
        same model (phenotype, kinship, cofactors) and calls the scan with its own slice of the SNPs (parallel.shard_range
        order); the n^3 set-up product is split over the ranks and the per-SNP results come back covering all SNPs
        (parallel.scan_sharded).  m_total = total number of SNPs when the slices follow parallel.shard_range.  Not set: a
        model never communicates, whatever process group exists (ranks may then fit different models independently)."""
        self.shard = {'group': None if group == 'world' else group, 'm_total': m_total}

    # ------------------------------------------------------------------------------------------
    def add_random_effect(self, cov_matrix=None, effect_type='normal'):
        """linear_models.py:577-580: stores kinship.scale_k(cov_matrix)."""
        if effect_type != 'normal':
            raise Exception('Currently, only Normal random effects are allowed.')
        own = False
        if isinstance(cov_matrix, DeviceMatrix):
            src = cov_matrix
        elif isinstance(cov_matrix, LazyHostArray) and cov_matrix._rows is None:
            src = cov_matrix.dev
        else:
            src = self.ctx.lookup_resident(cov_matrix)        # the kinship this context just handed out: still in HBM
            if src is None:
                src = DeviceMatrix.from_host(self.ctx, np.asarray(cov_matrix, dtype=np.float64))
                own = True
        if src.shape != (self.n, self.n):
            raise ValueError('kinship must be %d x %d' % (self.n, self.n))
        K = self.ctx.scale_k_copy(src)                        # :580 -- the caller's matrix is never modified
        if own:
            src.free()
        self._K_dev.append(K)
        self.random_effects.append((effect_type, LazyHostArray(K)))
        self._invalidate()

    def set_random_effect(self, cov_matrix_list, effect_types=None):
        """linear_models.py:583-586."""
        self.random_effects = [('normal', _LazyIdentity(self.n))]
        self._K_dev = []
        for cov_matrix in cov_matrix_list:
            self.add_random_effect(cov_matrix=kinship.scale_k(np.asarray(cov_matrix), ctx=self.ctx))

    def _K_device(self, K=None):
        if _is_none(K):
            return self._K_dev[0]
        return self.ctx.to_device(K)

    # ------------------------------------------------------------------------------------------
    def _get_eigen_L_(self, K=None, dtype='single'):
        """linear_models.py:589-596: eigh(K); 'vectors' holds the eigenvectors as rows."""
        Kd = self._K_device(K)
        U = Kd.copy()
        w = self.ctx.syevd(U)
        return EigenDict(values=w, vectors=LazyHostArray(U))

    def _get_eigen_R_(self, X=None, K=None, hat_matrix=None, dtype='single'):
        """linear_models.py:600-615: eigh(S(K+I)S), S = I - X(X'X)^+X'; drop the q smallest; values - 1."""
        if _is_none(X):
            X = self.X
        X = np.asarray(X, dtype=np.float64)
        q = X.shape[1]
        ctx = self.ctx
        Kd = self._K_device(K)
        A = Kd.copy()
        ctx.add_diag(A, 1.0)                                  # K + I   (:610)
        # S A S = A - X(G B') - (B G)X' + X(G C G)X' with B = A X, C = X'B, G = (X'X)^+  -- rank-q updates
        G = np.linalg.pinv(X.T @ X)                           # :605
        Xd = DeviceMatrix.from_host(ctx, X)
        B = ctx.gemm(A, Xd).download()                        # n x q
        Cm = X.T @ B
        left = DeviceMatrix.from_host(ctx, np.hstack([-X, -(B @ G), X @ (G @ Cm @ G)]))       # n x 3q
        right = DeviceMatrix.from_host(ctx, np.hstack([B @ G.T, X, X]))                       # n x 3q
        ctx.gemm(left, right, A, tb=True, beta=1.0)           # A += left right'
        for t in (Xd, left, right):
            t.free()
        w = ctx.syevd(A)
        eig_values = w[q:] - 1                                # :614
        return EigenDict(values=eig_values, vectors=LazyHostArray(A, rows=(q, self.n)), _q=q)

    def _rell_(self, delta, eig_vals, sq_etas):
        """linear_models.py:618-623."""
        num_eig_vals = len(eig_vals)
        c_1 = 0.5 * num_eig_vals * (np.log(num_eig_vals / (2.0 * np.pi)) - 1)
        v = eig_vals + delta
        res = c_1 - 0.5 * (num_eig_vals * np.log(np.sum(np.asarray(sq_etas).flatten() / v)) + np.sum(np.log(v)))
        return res

    def _redll_(self, delta, eig_vals, sq_etas):
        """linear_models.py:626-631."""
        num_eig_vals = len(eig_vals)
        v1 = eig_vals + delta
        v2 = np.asarray(sq_etas).flatten() / v1
        res = (num_eig_vals * np.sum(v2 / v1) / np.sum(v2) - np.sum(1.0 / v1))
        return res

    # ------------------------------------------------------------------------------------------
    def get_REML(self, ngrids=100, llim=-10, ulim=10, esp=1e-6, eig_L=None, eig_R=None):
        """
        Get REML estimates for the effect sizes, as well as the random effect contributions.
        This is EMMA (linear_models.py:653-668).
        """
        if not eig_L:
            eig_L = self._get_eigen_L_(None)
        res = self.get_estimates(eig_L, ngrids=ngrids, llim=llim, ulim=ulim, esp=esp, method='REML', eig_R=eig_R)
        res['eig_L'] = eig_L
        return res

    def get_ML(self, ngrids=100, llim=-10, ulim=10, esp=1e-6, eig_L=None, eig_R=None, H=None, H_inv=None, H_sqrt_inv=None,
               dtype='single'):
        """
        Get ML estimates for the effect sizes, as well as the random effect contributions (linear_models.py:672-696).
        H is the (full) covariance matrix, which speeds up calculations if it's available.
        """
        if _is_none(H):
            if not eig_L:
                eig_L = self._get_eigen_L_(None)
            return self.get_estimates(eig_L, ngrids=ngrids, llim=llim, ulim=ulim, esp=esp, method='ML', eig_R=eig_R)
        # the variance matrix is given: the likelihood from its eigenvalues and the transformed fit (:685-695)
        ctx = self.ctx
        Hd = ctx.to_device(H).copy()
        evals = ctx.syevd(Hd)
        Hd.free()
        Hs = ctx.to_device(H_sqrt_inv)
        XY = DeviceMatrix.from_host(ctx, np.hstack([self.X, self.Y]))
        t = ctx.gemm(Hs, XY).download()
        XY.free()
        q = self.X.shape[1]
        X_t, Y_t = t[:, :q], t[:, q:]
        (betas, mahalanobis_rss, rank, hs) = np.linalg.lstsq(X_t, Y_t, rcond=None)
        assert np.size(mahalanobis_rss) > 0, 'WTF?'
        rss = float(np.sum((self.Y - self.X @ betas) ** 2))
        n = Y_t.shape[0]
        ll = -0.5 * (n * np.log(2 * np.pi) + np.sum(np.log(evals)) + mahalanobis_rss)
        return {'ll': ll, 'rss': rss, 'mahalanobis_rss': mahalanobis_rss}

    def _etas(self, eig_R, Y):
        """etas = eig_R['vectors'] * Y (:794) with the eigenbasis resident in HBM."""
        ctx = self.ctx
        vec = eig_R['vectors']
        Yd = DeviceMatrix.from_host(ctx, Y)
        if isinstance(vec, LazyHostArray):
            full = ctx.gemm(vec.dev, Yd).download()
            r = vec._rows
            etas = full if r is None else full[r[0]:r[1]]
        else:
            Ud = DeviceMatrix.from_host(ctx, np.asarray(vec, dtype=np.float64))
            etas = ctx.gemm(Ud, Yd).download()
            Ud.free()
        Yd.free()
        return etas

    def get_estimates(self, eig_L, K=None, xs=None, ngrids=50, llim=-10, ulim=10, esp=1e-6,
                      return_pvalue=False, return_f_stat=False, method='REML', verbose=False,
                      dtype='single', eig_R=None, rss_0=None):
        """
        Get REML estimates for the effect sizes, as well as the random effect contributions, using the
        EMMA algorithm (Kang et al., Genetics, 2008)  -- linear_models.py:771-927, REML branch.
        """
        if method not in ('REML', 'ML'):
            raise Exception("method must be 'REML' or 'ML'")
        if xs is not None:
            xs = _col(xs, self.n)
            X = np.hstack([self.X, xs])
        else:
            X = self.X
        q = X.shape[1]
        n = self.n
        p = n - q
        m = ngrids + 1
        log_deltas = (np.arange(m, dtype=np.float64) / ngrids) * (ulim - llim) + llim      # :796
        deltas = np.exp(log_deltas)
        assert len(deltas) == m, 'Number of deltas is incorrect.'

        if method == 'ML' or xs is not None:
            # The reference eigendecomposes S(K+I)S for THIS X (:787-788: one n x n eigh per call, i.e. per tested SNP in
            # expedited_REML_t_test) only to evaluate four sums that are functions of eig_L and the rotated columns alone
            # (csrc/emma.cuh): no eig_R here.  REML without xs keeps the eig_R form below, which mirrors :789-810 literally.
            if K is not None:
                raise NotImplementedError('get_estimates(K=...) with xs / ML: pass the eig_L of that K instead')
            UL = self.ctx.to_device(eig_L['vectors'])
            r = self.ctx.emma(UL, eig_L['values'], X, self.Y, deltas=deltas, esp=esp, method=method, want_grid=True)
            opt_delta = float(r['delta'][0])
            opt_ll = float(r['max_ll'][0])
            opt_vg = float(r['vg'][0])
            self._last_reml = {'lls': r['lls'][0], 'dlls': r['dlls'][0], 'deltas': deltas, 'flags': None, 'eig_R': None}
        else:
            if not eig_R:
                if self._eig_R_cache is not None and _is_none(K):
                    eig_R = self._eig_R_cache
                else:
                    eig_R = self._get_eigen_R_(X=X, K=K)
                    if _is_none(K):
                        self._eig_R_cache = eig_R
            etas = self._etas(eig_R, self.Y)                                       # :794
            sq_etas = etas * etas
            eig_vals = np.array(eig_R['values'], dtype=np.float64)
            assert len(eig_vals) == p, 'Number of eigenvalues is incorrect.'
            r = self.ctx.reml(eig_vals, sq_etas.T, deltas, esp)                    # :802-891 on the device
            opt_delta = float(r['delta'][0])
            opt_ll = float(r['ll'][0])
            # :894-896 -- the reference divides a (p,1) by a (p,) array: a p x p outer quotient
            opt_vg = float(np.sum(sq_etas) * np.sum(1.0 / (eig_vals + opt_delta)) / p)
            self._last_reml = {'lls': r['lls'][0], 'dlls': r['dlls'][0], 'deltas': deltas, 'flags': int(r['flags'][0]),
                               'eig_R': eig_R}
        opt_ve = opt_vg * opt_delta

        # H_sqrt_inv = diag(1/sqrt(eig_L.values + delta)) eig_L.vectors   (:898) -- kept as its two factors
        ctx = self.ctx
        UL = ctx.to_device(eig_L['vectors'])
        H = LazyScaledRows(UL, 1.0 / np.sqrt(np.asarray(eig_L['values'], dtype=np.float64) + opt_delta))
        t = H.times(np.hstack([X, self.Y]))
        X_t, Y_t = t[:, :q], t[:, q:]
        (beta_est, mahalanobis_rss, rank, sigma) = np.linalg.lstsq(X_t, Y_t, rcond=None)
        if np.size(mahalanobis_rss) == 0:
            mahalanobis_rss = np.array([np.sum((Y_t - X_t @ beta_est) ** 2)])
        x_beta = X @ beta_est
        residuals = self.Y - x_beta
        rss = residuals.T @ residuals
        res_dict = {'max_ll': opt_ll, 'delta': opt_delta, 'beta': beta_est, 've': opt_ve, 'vg': opt_vg,
                    'rss': rss, 'mahalanobis_rss': mahalanobis_rss, 'H_sqrt_inv': H,
                    'pseudo_heritability': 1.0 / (1 + opt_delta)}
        if xs is not None and return_f_stat:
            h0_X = X_t[:, :self.X.shape[1]]
            (h0_betas, h0_rss, h0_rank, h0_s) = np.linalg.lstsq(h0_X, Y_t, rcond=None)
            f_stat = (h0_rss / mahalanobis_rss - 1) * p / xs.shape[1]
            res_dict['var_perc'] = 1.0 - mahalanobis_rss / h0_rss
            res_dict['f_stat'] = float(np.asarray(f_stat).reshape(-1)[0])
        if return_pvalue:
            p_val = ctx.f_sf(np.array([res_dict['f_stat']]), xs.shape[1], p)
            res_dict['p_val'] = float(p_val[0])
        return res_dict

    def expedited_REML_t_test(self, snps, ngrids=50, llim=-4, ulim=10, esp=1e-6, verbose=True, eig_L=None, _resident_rows=None):
        """
        Single SNP analysis, i.e. EMMA (linear_models.py:931-968).  The reference calls get_estimates once per SNP, each call
        eigendecomposing S(K+I)S for X = [X0, snp] (:787-788); here ALL the SNPs are fitted in one batch from eig_L alone
        (mmg_emma_f64, csrc/emma.cuh): one rotation GEMM, then the delta grid, the secant refinement and the GLS fit per SNP in
        two kernel launches.  Same outputs; identical in exact arithmetic (FP64 here, float32 grid in the reference).
        `_resident_rows`: rows of the resident genotype block to refine instead of `snps` (the scan's own top hits).
        """
        assert len(self.random_effects) == 2, "Expedited REMLE only works when we have exactly two random effects."
        if _is_none(eig_L):
            eig_L = self._get_eigen_L_(None)
        deltas = np.exp((np.arange(ngrids + 1, dtype=np.float64) / ngrids) * (ulim - llim) + llim)
        UL = self.ctx.to_device(eig_L['vectors'])
        if _resident_rows is not None:
            r = self.ctx.emma(UL, eig_L['values'], self.X, self.Y, snp_rows=_resident_rows, deltas=deltas, esp=esp)
        else:
            if len(snps) == 0:
                return {'ps': np.empty(0), 'f_stats': np.empty(0), 'vgs': np.empty(0), 'ves': np.empty(0), 'var_perc': np.empty(0),
                        'max_lls': np.empty(0), 'betas': [], 'rss': np.empty(0)}
            xs = np.asarray([np.asarray(x, dtype=np.float64).reshape(-1) for x in snps])
            r = self.ctx.emma(UL, eig_L['values'], self.X, self.Y, xs=xs, deltas=deltas, esp=esp)
        return {'ps': r['p_val'], 'f_stats': r['f_stat'], 'vgs': r['vg'], 'ves': r['ve'], 'var_perc': r['var_perc'],
                'max_lls': r['max_ll'], 'betas': [list(map(float, b)) for b in r['betas']], 'rss': r['rss']}

    # ------------------------------------------------------------------------------------------
    def emmax_f_test(self, snps, snp_priors=None, Z=None, with_betas=False, method='REML',
                     eig_L=None, eig_R=None, emma_num=100):
        """
        EMMAX implementation, single SNPs (linear_models.py:1233-1267).
        """
        if not eig_L:
            _say('Calculating the eigenvalues of K')
            s0 = time.time()
            eig_L = self._get_eigen_L_()
            _say('Done.\nTook %0.2f seconds' % (time.time() - s0))
        if not eig_R:
            _say("Calculating the eigenvalues of S(K+I)S where S = I-X(X'X)^-1X'")
            s0 = time.time()
            eig_R = self._get_eigen_R_(X=self.X)
            _say('Done\nTook %0.2f seconds' % (time.time() - s0))

        _say('Getting variance estimates')
        s0 = time.time()
        res = self.get_estimates(eig_L, method=method, eig_R=eig_R)
        _say('Done.\nTook %0.2f seconds' % (time.time() - s0))
        _say('pseudo_heritability:', res['pseudo_heritability'])

        s0 = time.time()
        r = self._emmax_f_test_(snps, res['H_sqrt_inv'], snp_priors=snp_priors, Z=Z, with_betas=with_betas,
                                emma_num=emma_num, eig_L=eig_L)
        _say('Took %0.2f seconds' % (time.time() - s0))
        r['pseudo_heritability'] = res['pseudo_heritability']
        r['ve'] = res['ve']
        r['vg'] = res['vg']
        r['max_ll'] = res['max_ll']
        return r

    def _null_fit(self, H, Z=None, project=True):
        """Set-up of _emmax_f_test_ (linear_models.py:1290-1306): null GLS fit in the rotated space and
        the rotation R = M' = (I - QQ')H (or H with no projection), resident in HBM.  H: a DeviceMatrix, or the factored
        LazyScaledRows that get_estimates returns -- then R is formed from the eigenbasis in one pass."""
        ctx = self.ctx
        n = self.n
        q0 = self.X.shape[1]
        factored = isinstance(H, LazyScaledRows) and Z is None
        if factored:
            t = H.times(np.hstack([self.X, self.Y]))
        else:
            if isinstance(H, LazyScaledRows):
                H = H.dev
            XY = DeviceMatrix.from_host(ctx, np.hstack([self.X, self.Y]))
            t = ctx.gemm(H, XY).download()
            XY.free()
        h0_X, Y = t[:, :q0], t[:, q0:]
        (h0_betas, h0_rss, h0_rank, h0_s) = np.linalg.lstsq(h0_X, Y, rcond=None)
        Yres = Y - h0_X @ h0_betas
        if np.size(h0_rss) == 0:
            h0_rss = np.array([np.sum(Yres ** 2)])
        if factored:
            Q = np.linalg.qr(h0_X)[0] if project else None        # :1300
            Rm = ctx.rotation(H.U, H.d, Q)                        # :1303 (transposed) / :1306
            return {'h0_X': h0_X, 'h0_betas': h0_betas, 'h0_rss': h0_rss, 'Yres': Yres, 'R': Rm, 'owned': True}
        Hz = H
        if Z is not None:
            Zd = DeviceMatrix.from_host(ctx, np.asarray(Z, dtype=np.float64))
            Hz = ctx.gemm(H, Zd)                              # :1297  H <- H Z
            Zd.free()
        if project:
            (Q, Rq) = np.linalg.qr(h0_X)                      # :1300
            Qd = DeviceMatrix.from_host(ctx, Q)
            QtH = ctx.gemm(Qd, Hz, ta=True)                   # q0 x n
            Rm = Hz.copy() if Hz is H else Hz
            ctx.gemm(Qd, QtH, Rm, alpha=-1.0, beta=1.0)       # R = H - Q (Q'H)      (:1303, transposed)
            Qd.free()
            QtH.free()
        else:
            Rm = Hz                                           # :1306  M = H'
        return {'h0_X': h0_X, 'h0_betas': h0_betas, 'h0_rss': h0_rss, 'Yres': Yres, 'R': Rm, 'owned': Rm is not H}

    def _emmax_f_test_(self, snps, H_sqrt_inv, snp_priors=None, verbose=True, return_transformed_snps=False,
                       Z=None, with_betas=False, emma_num=100, eig_L=None, **kwargs):
        """
        EMMAX implementation, single SNPs (linear_models.py:1272-1380).  The two hot loops (:1316-1339) --
        the rotation GEMM and the per-SNP least squares -- and the F / p-value epilogue (:1345-1349) run
        as one fused kernel over the resident genotypes.
        """
        ctx = self.ctx
        q = 1  # Single SNP is being tested
        p = len(self.X.T) + q
        n = self.n
        n_p = n - p
        xs_real = _lib.real_valued(snps)          # imputed dosages: FP64 tensor-core scan of the rows themselves (:1317 takes any numeric row)
        if xs_real is not None:
            num_snps, n_lines = xs_real.shape
            scan = lambda R_, V_, h_, np_, **kw: ctx.emmax_scan_rows(xs_real, R_, V_, h_, np_, **kw)
        else:
            num_snps, n_lines = ctx.ensure_snps(snps)
            scan = ctx.emmax_scan
        H = H_sqrt_inv if isinstance(H_sqrt_inv, LazyScaledRows) else ctx.to_device(H_sqrt_inv)
        nf = self._null_fit(H, Z=Z, project=not with_betas)
        h0_rss = nf['h0_rss']
        h0_rss_f = float(np.asarray(h0_rss).reshape(-1)[0])
        h0_betas = list(map(float, list(np.asarray(nf['h0_betas']).reshape(-1))))
        Rm = nf['R']
        if Rm.shape[1] != n_lines:
            raise ValueError('SNP length %d does not match the model (%d)' % (n_lines, Rm.shape[1]))
        impl = kwargs.get('impl', self.scan_impl)

        if not with_betas:
            int8_scan = impl == 'tcgen05' or (impl in ('auto', None) and os.environ.get('MMG_SCAN_IMPL', 'tcgen05') == 'tcgen05')
            if self.shard is not None and xs_real is None:
                # one process per GPU, explicitly requested (set_sharding): R'R is formed once across the ranks instead of once
                # per rank, every rank scans its SNP slice, the outputs are all-gathered (parallel.py)
                from . import parallel
                if not int8_scan:
                    raise ValueError("the sharded scan runs on the int8 tensor-core path (scan_impl='tcgen05')")
                group, m_total = self.shard['group'], self.shard['m_total']
                if parallel.world_size(group) > 1:
                    if emma_num > 0 or return_transformed_snps or snp_priors is not None:
                        raise ValueError('the sharded scan returns the per-SNP statistics only (emma_num=0, no t_snps / priors): '
                                         'refine the top hits on one rank with expedited_REML_t_test')
                    parallel.check_same_model(ctx, [h0_rss_f, float(n), float(np.sum(np.abs(nf['Yres']))), float(Rm.shape[0])], group)
                    ctx.scan_prepass_begin(Rm, nf['Yres'])         # v = R'y~ and the linear terms: side stream, under the next two
                    A, a_err = parallel.quad_form_sharded(ctx, Rm, group)
                    out = parallel.scan_sharded(ctx, A, a_err, None, h0_rss_f, n_p, m_total=m_total, group=group)
                    A.free()
                    num_snps = len(out['ps'])
                else:
                    out = scan(Rm, nf['Yres'].reshape(1, -1), h0_rss_f, n_p, impl=impl)
            else:
                out = scan(Rm, nf['Yres'].reshape(1, -1), h0_rss_f, n_p, impl=impl)
            p_vals, f_stats, rss_list, var_perc = out['ps'], out['f_stats'], out['rss'], out['var_perc']
        else:
            # lstsq([h0_X, x~], Y) (:1323) through its normal equations: the scan supplies xx = x~.x~, xy = x~.Yres, b = x~.h0_X;
            # the (q0+1)x(q0+1) solve is a Schur complement per SNP
            h0_X, Yres = nf['h0_X'], nf['Yres']
            if xs_real is None:
                # ... evaluated on the device, with F and p (betas_finish_kernel): no statistic is computed on the host
                out = ctx.emmax_scan_betas(Rm, Yres, h0_X, h0_betas, h0_rss_f, n_p, impl=impl)
                p_vals, f_stats, rss_list, var_perc = out['ps'], out['f_stats'], out['rss'], out['var_perc']
                kept = np.isnan(out['betas'][:, -1])
                betas_list = out['betas'].tolist()
                for i in np.flatnonzero(kept):
                    betas_list[i] = h0_betas
            else:
                # real-valued rows (dosages): the moments come back and the small solve runs vectorised on the host
                V = np.vstack([Yres.T, h0_X.T])
                out = scan(Rm, V, h0_rss_f, n_p, impl=impl, want_dots=True, want_stats=False)
                xx, dots = out['xx'], out['dots']
                xy, b = dots[:, 0], dots[:, 1:]
                A = h0_X.T @ h0_X
                c0 = (h0_X.T @ Yres).reshape(-1)
                Ainv = np.linalg.inv(A)
                Ab = b @ Ainv.T                                   # rows: A^-1 b_s
                s = xx - np.einsum('ij,ij->i', b, Ab)
                ok = s > 1e-12 * np.maximum(xx, 1e-300)
                s_safe = np.where(ok, s, 1.0)
                beta_x = (xy - Ab @ c0) / s_safe
                beta_0 = (Ainv @ c0)[None, :] - Ab * beta_x[:, None]
                yy = float(np.sum(Yres ** 2))
                rss_full = yy - (beta_0 @ c0 + beta_x * xy)
                good = ok & (rss_full != 0)
                rss_list = np.where(good, rss_full, h0_rss_f)
                betas_arr = np.hstack([beta_0, beta_x[:, None]])
                betas_list = [list(map(float, row)) if g else h0_betas for row, g in zip(betas_arr, good)]
                rss_ratio = h0_rss_f / rss_list
                var_perc = 1 - 1 / rss_ratio
                f_stats = (rss_ratio - 1) * n_p / float(q)
                p_vals = ctx.f_sf(f_stats, q, n_p)

        res_d = {'ps': p_vals, 'f_stats': f_stats, 'rss': rss_list, 'var_perc': var_perc,
                 'h0_rss': h0_rss, 'h0_betas': h0_betas}
        if with_betas:
            res_d['betas'] = betas_list
        if return_transformed_snps:
            res_d['t_snps'] = self._transformed_snps(snps, Rm)
        if snp_priors is not None:
            snp_priors = np.array(snp_priors)
            log_bfs = np.where(rss_list != h0_rss_f, np.log(h0_rss_f) - np.log(rss_list), 0.0)      # :1335
            bfs = np.exp((log_bfs * n - np.log(n)) * 1 / 2)                                          # :1358
            res_d['bfs'] = bfs
            pos = bfs * snp_priors / (1 - snp_priors)
            res_d['pos'] = pos
            res_d['ppas'] = pos / (1 + pos)
        if nf['owned']:
            Rm.free()

        if emma_num > 0:                                                                            # :1365-1377
            pval_indices = sorted(zip(res_d['ps'], range(num_snps)))[:emma_num]
            _say('Updating p-values using EMMA for the smallest %d p-values.' % len(pval_indices))
            l = list(map(list, zip(*pval_indices)))
            # the top hits are refined straight from the resident genotype block (the scan just used it): no re-upload
            if xs_real is not None:
                top_emma_res = self.expedited_REML_t_test(xs_real[np.asarray(l[1], dtype=np.int64)], eig_L=eig_L)
            else:
                top_emma_res = self.expedited_REML_t_test(None, eig_L=eig_L, _resident_rows=np.asarray(l[1], dtype=np.int64))
            for pi, pv, f, r, v in zip(l[1], top_emma_res['ps'], top_emma_res['f_stats'],
                                       top_emma_res['rss'], top_emma_res['var_perc']):
                res_d['ps'][pi] = pv
                res_d['f_stats'][pi] = f
                res_d['rss'][pi] = r
                res_d['var_perc'][pi] = v
        return res_d

    def _transformed_snps(self, snps, Rm, chunk=4096):
        """t_snps of :1320-1321: the rotated SNPs x~ = R x, materialised (small m only)."""
        ctx = self.ctx
        out = []
        m = len(snps)
        for s0 in range(0, m, chunk):
            xc = DeviceMatrix.from_host(ctx, np.asarray(snps[s0:s0 + chunk], dtype=np.float64))
            t = ctx.gemm(xc, Rm, tb=True).download()
            xc.free()
            out.extend(list(t))
        return out

    # ------------------------------------------------------------------------------------------
    def emmax_GxT_f_test(self, snps, E, Z=None, with_betas=False, method='REML', eig_L=None, eig_R=None):
        """
        EMMAX with a genotype x environment term, single SNPs (linear_models.py:1383-1416).
        """
        if not eig_L:
            eig_L = self._get_eigen_L_()
        if not eig_R:
            eig_R = self._get_eigen_R_(X=self.X)
        res = self.get_estimates(eig_L, method=method, eig_R=eig_R)
        r = self._emmax_GxT_f_test_(snps, res['H_sqrt_inv'], E, Z, with_betas=with_betas, eig_L=eig_L)
        r['pseudo_heritability'] = res['pseudo_heritability']
        r['ve'] = res['ve']
        r['vg'] = res['vg']
        r['max_ll'] = res['max_ll']
        return r

    def _moments(self, x, Rm, V, h0_rss_f, n_p, impl):
        """x~.x~ and x~.V[k] of every row of the genotype block x through the fused scan kernels (int rows: resident block;
        real-valued rows: FP64 tensor-core path)."""
        ctx = self.ctx
        xr = _lib.real_valued(x)
        if xr is not None:
            out = ctx.emmax_scan_rows(xr, Rm, V, h0_rss_f, n_p, want_dots=True, want_stats=False)
        else:
            ctx.ensure_snps(x)
            out = ctx.emmax_scan(Rm, V, h0_rss_f, n_p, impl=impl, want_dots=True, want_stats=False)
        return np.array(out['xx']), np.array(out['dots'])

    def _emmax_GxT_f_test_(self, snps, H_sqrt_inv, T, Z, verbose=True, **kwargs):
        """
        EMMAX G and G x T tests, single SNPs (linear_models.py:1422-1514): per SNP the genetic model [h0_X, x~] and the full
        model [h0_X, x~, t~] with x~ = H x, t~ = H (x o T) (:1447-1449, M = H', no Q projection).  The reference solves two
        lstsq per SNP; here the moments of both models come from three passes of the fused scan kernel -- over the blocks x,
        x o T and x o (1 + T): x~.x~, t~.t~ and, by polarisation, x~.t~ = (s~.s~ - x~.x~ - t~.t~) / 2 -- together with the dot
        products against the residual phenotype and h0_X; the (q0+1) and (q0+2) normal equations are then Schur complements,
        vectorised over the SNPs.  Returns {'g_res', 'gt_res', 'gt_g_res'} with the reference's keys.
        """
        ctx = self.ctx
        n = self.n
        impl = kwargs.get('impl', 'dmma' if self.scan_impl in ('auto', None) else self.scan_impl)   # differences of moments: FP64 pipe by default
        H = H_sqrt_inv if isinstance(H_sqrt_inv, LazyScaledRows) else ctx.to_device(H_sqrt_inv)
        nf = self._null_fit(H, Z=None, project=False)                       # :1437-1441, M = H' (:1446)
        h0_X, Yres, Rm = nf['h0_X'], nf['Yres'], nf['R']
        h0_rss = nf['h0_rss']
        h0_rss_f = float(np.asarray(h0_rss).reshape(-1)[0])
        h0_betas = list(map(float, list(np.asarray(nf['h0_betas']).reshape(-1))))
        q0 = h0_X.shape[1]
        T_flat = np.array(T, dtype=np.float64).flatten()
        x = np.asarray(snps)
        if x.ndim != 2:
            x = np.asarray([np.asarray(r) for r in snps])
        if Z is not None:
            x = x @ np.asarray(Z).T                                          # :1452
        if x.shape[1] != n or T_flat.shape[0] != n:
            raise ValueError('SNPs / environment vector do not match the model (%d individuals)' % n)
        num_snps = x.shape[0]
        Ti = np.rint(T_flat)
        if np.array_equal(Ti, T_flat) and x.dtype.kind in 'iub':
            Tm = Ti.astype(np.int64)
            xt, xs_ = x * Tm, x * (1 + Tm)
        else:
            xt, xs_ = x * T_flat, x * (1.0 + T_flat)
        V = np.vstack([Yres.T, h0_X.T])
        xx, dg = self._moments(x, Rm, V, h0_rss_f, n - q0 - 1, impl)
        tt, dt = self._moments(xt, Rm, V, h0_rss_f, n - q0 - 1, impl)
        ss, _ = self._moments(xs_, Rm, V, h0_rss_f, n - q0 - 1, impl)
        ctx.invalidate_snps()
        if nf['owned']:
            Rm.free()
        gt = 0.5 * (ss - xx - tt)
        xy, b = dg[:, 0], dg[:, 1:]
        ty, bt = dt[:, 0], dt[:, 1:]
        A = h0_X.T @ h0_X
        c0 = (h0_X.T @ Yres).reshape(-1)
        Ainv = np.linalg.inv(A)
        a0 = Ainv @ c0
        Ab, Abt = b @ Ainv.T, bt @ Ainv.T
        yy = float(np.sum(Yres ** 2))
        tiny = 1e-12
        # genetic model (:1457-1460)
        s11 = xx - np.einsum('ij,ij->i', b, Ab)
        r1 = xy - Ab @ c0
        ok_g = s11 > tiny * np.maximum(xx, 1e-300)
        s11s = np.where(ok_g, s11, 1.0)
        bx = r1 / s11s
        b0 = a0[None, :] - Ab * bx[:, None]
        rss_g = yy - (b0 @ c0 + bx * xy)
        good_g = ok_g & (rss_g != 0)
        rss_g_list = np.where(good_g, rss_g, h0_rss_f)
        betas_g = np.hstack([b0, bx[:, None]])
        # full model (:1462-1465), fitted only where the genetic model was (:1458)
        s12 = gt - np.einsum('ij,ij->i', b, Abt)
        s22 = tt - np.einsum('ij,ij->i', bt, Abt)
        r2 = ty - Abt @ c0
        det = s11 * s22 - s12 * s12
        ok_gt = good_g & (s22 > tiny * np.maximum(tt, 1e-300)) & (det > 1e-10 * np.abs(s11 * s22))
        dets = np.where(ok_gt, det, 1.0)
        g1 = (s22 * r1 - s12 * r2) / dets
        g2 = (s11 * r2 - s12 * r1) / dets
        g0 = a0[None, :] - Ab * g1[:, None] - Abt * g2[:, None]
        rss_gt = yy - (g0 @ c0 + g1 * xy + g2 * ty)
        good_gt = ok_gt & (rss_gt != 0)
        rss_gt_list = np.where(good_gt, rss_gt, h0_rss_f)
        betas_gt = np.hstack([g0, g1[:, None], g2[:, None]])
        betas_g_list = [list(map(float, row)) if g else h0_betas for row, g in zip(betas_g, good_g)]
        betas_gt_list = [list(map(float, row)) if g else h0_betas for row, g in zip(betas_gt, good_gt)]

        def f_test(num, den, q):
            n_p = n - (q0 + q)
            ratio = num / den
            f = (ratio - 1) * n_p / float(q)
            return ctx.f_sf(np.maximum(f, 0.0), q, n_p), f, 1 - 1 / ratio

        ps, f, vp = f_test(h0_rss_f, rss_g_list, 1)                         # :1478-1486
        g_res_d = {'ps': ps, 'f_stats': f, 'rss': rss_g_list, 'var_perc': vp, 'h0_rss': h0_rss, 'h0_betas': h0_betas,
                   'betas': betas_g_list}
        ps, f, vp = f_test(h0_rss_f, rss_gt_list, 2)                        # :1490-1498
        gt_res_d = {'ps': ps, 'f_stats': f, 'rss': rss_gt_list, 'var_perc': vp, 'betas': betas_gt_list}
        ps, f, vp = f_test(rss_g_list, rss_gt_list, 1)                      # :1502-1511 (p = len(X') + 1, as the reference has it)
        gt_g_res_d = {'ps': ps, 'f_stats': f, 'var_perc': vp}
        return {'g_res': g_res_d, 'gt_res': gt_res_d, 'gt_g_res': gt_g_res_d}

    # ------------------------------------------------------------------------------------------
    def _emmax_permutations_(self, snps, K, H_sqrt_inv, num_perm=100):
        """
        EMMAX permutation test, single SNPs (linear_models.py:1125-1175).  Returns the list of min p-values
        and max F statistics.  Reference quirks kept: self.Y is mean-centred in place (:1140); the null fit
        is subtracted twice (:1144,:1147); SNPs are centred and rotated by H' with no Q projection
        (:1159-1160); the permuted phenotypes come from cumulative in-place np.random.shuffle calls on the
        global legacy RNG (:1151-1154), so seeding np.random reproduces the reference's Ys.
        """
        ctx = self.ctx
        q = 1
        p = len(self.X.T) + q
        n = self.n
        n_p = n - p
        if self.X.shape[1] != 1:
            raise ValueError('the reference (:1147) only type-checks with a single fixed effect')
        self.Y = self.Y - np.mean(self.Y)                                      # :1140
        H = ctx.to_device(H_sqrt_inv)
        XY = DeviceMatrix.from_host(ctx, np.hstack([self.X, self.Y]))
        t = ctx.gemm(H, XY).download()
        XY.free()
        h0_X, Y = t[:, :1], t[:, 1:]
        (h0_betas, h0_rss, h0_rank, h0_s) = np.linalg.lstsq(h0_X, Y, rcond=None)
        Y = Y - h0_X @ h0_betas                                                # :1144
        Y = Y - h0_X * float(h0_betas[0, 0])                                   # :1147
        Ys = np.zeros((n, num_perm))
        y_col = Y[:, 0]                                                        # 1-D view of the (n, 1) column
        for perm_i in range(num_perm):
            # :1153 `sp.random.shuffle(Y)`: shuffling the 1-D view draws the same random_interval sequence from the legacy global
            # RNG and produces the same permutation as shuffling the (n, 1) array, without its row-by-row buffer swaps (5.7 ms -> 0.09 ms
            # per permutation at n = 5000; tests/test_host_logic.py pins the equivalence)
            np.random.shuffle(y_col)
            Ys[:, perm_i] = y_col
        h0_rss_f = float(np.asarray(h0_rss).reshape(-1)[0])

        num_snps, n_lines = ctx.ensure_snps(snps)
        # x~_c . Ys_p = x_c . (H' Ys_p): W' = Ys' H  ([P x n])
        Ysd = DeviceMatrix.from_host(ctx, Ys)
        Wt = ctx.gemm(Ysd, H, ta=True)
        Ysd.free()
        ratio = np.zeros(num_perm)
        ctx.emmax_perm_scan(H, Wt, ratio, centre=True, impl=getattr(self, 'perm_impl', 'auto'))
        Wt.free()
        ynorm = np.sum(Ys * Ys, axis=0)
        min_rss_list = np.minimum(np.repeat(h0_rss_f, num_perm), ynorm - ratio)      # :1156,:1164
        max_f_stats = ((h0_rss_f / min_rss_list) - 1.0) * n_p / float(q)             # :1171
        min_pvals = ctx.f_sf(max_f_stats, q, n_p)                                    # :1172
        return {'min_ps': min_pvals, 'max_f_stats': max_f_stats}


# ----------------------------------------------------------------------------------------------
def get_emma_reml_estimates(y, K, K2=None, cofactors=None, include_intercept=True):
    """linear_models.py:1690-1706 (single-kinship form)."""
    if K2 is not None:
        raise NotImplementedError('two-kinship models are outside the EMMAX hot path')
    lmm = LinearMixedModel(y)
    lmm.add_random_effect(K)
    if cofactors is not None:
        lmm.set_factors(cofactors, include_intercept=include_intercept)
    res = lmm.get_REML()
    H = np.asarray(res['H_sqrt_inv'])
    res['Y_t'] = H @ lmm.Y
    res['X_t'] = H @ lmm.X
    res['lmm'] = lmm
    return res


def emma(snps, phenotypes, K, cofactors=None):
    """Run EMMA (linear_models.py:1725-1745)."""
    lmm = LinearMixedModel(phenotypes)
    lmm.add_random_effect(K)
    if cofactors:
        for cofactor in cofactors:
            lmm.add_factor(cofactor)
    return lmm.expedited_REML_t_test(snps)


def emmax(snps, phenotypes, K, cofactors=None, Z=None, with_betas=False, emma_num=0, scan_impl='auto', ctx=None, shard=None):
    """
    Run EMMAX (linear_models.py:1790-1816).  `shard` (not in the reference): see LinearMixedModel.set_sharding.
    """
    lmm = LinearMixedModel(phenotypes, ctx=ctx, scan_impl=scan_impl, shard=shard)
    if Z is not None:
        Zm = np.asarray(Z, dtype=np.float64)
        Kh = np.asarray(K, dtype=np.float64)
        lmm.add_random_effect(Zm @ Kh @ Zm.T)
        if cofactors:
            for cofactor in cofactors:
                lmm.add_factor(Zm @ np.asarray(cofactor, dtype=np.float64))
    else:
        lmm.add_random_effect(K)
        if cofactors:
            for cofactor in cofactors:
                lmm.add_factor(cofactor)

    _say("Running EMMAX")
    s1 = time.time()
    res = lmm.emmax_f_test(snps, Z=Z, with_betas=with_betas, emma_num=emma_num)
    secs = time.time() - s1
    if secs > 60:
        mins = int(secs) // 60
        secs = secs - mins * 60
        _say('Took %d mins and %f seconds.' % (mins, secs))
    else:
        _say('Took %f seconds.' % (secs))
    return res


def emmax_w_two_env(snps, phenotypes, K, E, cofactors=None, Z=None, ctx=None):
    """
    Run EMMAX with environmental variables (linear_models.py:1749-1787).  Three p-values per SNP: the genetic model against
    the null ('g_res'), the full model (genetic + genotype x environment) against the null ('gt_res'), and the full model
    against the genetic model ('gt_g_res').  E: 0-1 (or real) vector distinguishing the environments; Z: incidence matrix
    for replicates.
    """
    lmm = LinearMixedModel(phenotypes, ctx=ctx)
    if Z is not None:
        Zm = np.asarray(Z, dtype=np.float64)
        lmm.add_random_effect(Zm @ np.asarray(K, dtype=np.float64) @ Zm.T)
        if cofactors:
            for cofactor in cofactors:
                lmm.add_factor(Zm @ np.asarray(cofactor, dtype=np.float64))
    else:
        lmm.add_random_effect(K)
        if cofactors:
            for cofactor in cofactors:
                lmm.add_factor(cofactor)
    _say("Running EMMAX w G and GxE tests")
    return lmm.emmax_GxT_f_test(snps, E=E, Z=Z)


def emmax_multi(snps, phenotypes, K, cofactors=None, ngrids=50, llim=-10, ulim=10, esp=1e-6, batch=None, ctx=None, shared=None):
    """
    EMMAX for T phenotypes measured on the same individuals (BASELINE.json configs[2]: "all phenotypes scanned
    against one kinship eigenbasis").  Equivalent to [emmax(snps, y, K, cofactors) for y in phenotypes]
    (linear_models.py:1790-1816, emma_num=0) with the work shared:
      * K is scaled and eigendecomposed once (eig_L and eig_R do not depend on the phenotype);
      * the REML grid + secant refinement of all T phenotypes is one launch pair (mmg_reml_f64, T rows);
      * the scan rotates every SNP ONCE, g = U_L x on the int8 tensor cores, for all phenotypes (mmg_emmax_scan_shared_f64):
        per phenotype only sum_k g_k^2 / (lambda_k + delta_t) (an FP64 tensor-core contraction, O(n) per SNP and phenotype) and
        the dot products with v_t = R_t' y~_t and the projected-out fixed effects remain -- one rotation + O(n T) per SNP
        instead of T rotations; the null fits need U_L [X, Y] (one skinny GEMM) and no per-phenotype n x n matrix.
    `shared=False` (default for T < 4) scans each phenotype with its own rotation R_t in one launch of the quadratic-form
    kernel instead.  `phenotypes` is a sequence of T length-n vectors.  Returns a list of T result dicts with emmax()'s keys.
    `batch` bounds the number of phenotypes per launch.
    """
    Y = np.asarray(phenotypes, dtype=np.float64)
    if Y.ndim != 2:
        raise ValueError('phenotypes must be a sequence of T equally long vectors')
    T, n = Y.shape
    base = LinearMixedModel(Y[0], ctx=ctx)
    ctx = base.ctx
    base.add_random_effect(K)
    if cofactors:
        for cofactor in cofactors:
            base.add_factor(cofactor)
    X = base.X
    q0 = X.shape[1]
    p = n - q0
    n_p = n - (q0 + 1)
    eig_L = base._get_eigen_L_()
    eig_R = base._get_eigen_R_(X=X)
    eig_vals = np.array(eig_R['values'], dtype=np.float64)
    eigL_vals = np.asarray(eig_L['values'], dtype=np.float64)
    etas = base._etas(eig_R, Y.T)                                             # p x T   (:794 for every phenotype)
    sq_etas = etas * etas
    m_g = ngrids + 1
    deltas = np.exp((np.arange(m_g, dtype=np.float64) / ngrids) * (ulim - llim) + llim)
    r = ctx.reml(eig_vals, sq_etas.T, deltas, esp)                            # :802-891, T phenotypes at once
    if _lib.real_valued(snps) is not None:
        raise TypeError('emmax_multi scans integer genotype codes (dosages: call emmax per phenotype)')
    num_snps, n_lines = ctx.ensure_snps(snps)
    if n_lines != n:
        raise ValueError('SNP length %d does not match the phenotypes (%d)' % (n_lines, n))
    UL = ctx.to_device(eig_L['vectors'])
    if shared is None:
        shared = T >= 4 and os.environ.get('MMG_MULTI_SHARED', '1') != '0'

    def meta_of(t, delta, h0_rss, h0_betas):
        vg = float(np.sum(sq_etas[:, t]) * np.sum(1.0 / (eig_vals + delta)) / p)          # :894-896
        return {'h0_rss': h0_rss, 'h0_betas': list(map(float, np.asarray(h0_betas).reshape(-1))),
                'pseudo_heritability': 1.0 / (1 + delta), 'vg': vg, 've': vg * delta, 'max_ll': float(r['ll'][t]), 'delta': delta}

    results = []
    if shared:
        XYd = DeviceMatrix.from_host(ctx, np.hstack([X, Y.T]))
        UXY = ctx.gemm(UL, XYd).download()                                   # U_L [X, Y]: H_t [X, y_t] = diag(d_t) of its columns (:1290)
        XYd.free()
        UX, UY = UXY[:, :q0], UXY[:, q0:]
        if batch is None:
            batch = max(1, min(T, int(2.0e9 / (8.0 * max(num_snps, 1)))))    # <= ~2 GB per output array
        for t0 in range(0, T, batch):
            ts = list(range(t0, min(T, t0 + batch)))
            Z = np.empty((n, len(ts) * (1 + q0)))
            W = np.empty((len(ts), n))
            h0, meta = [], []
            for i, t in enumerate(ts):
                delta = float(r['delta'][t])
                w = 1.0 / (eigL_vals + delta)
                d = np.sqrt(w)                                               # :898
                h0_X = d[:, None] * UX
                y_t = (d * UY[:, t]).reshape(-1, 1)
                (h0_betas, h0_rss, h0_rank, h0_s) = np.linalg.lstsq(h0_X, y_t, rcond=None)      # :1292
                Yres = y_t - h0_X @ h0_betas
                if np.size(h0_rss) == 0:
                    h0_rss = np.array([np.sum(Yres ** 2)])
                Q = np.linalg.qr(h0_X)[0]                                    # :1300
                Z[:, i * (1 + q0)] = d * Yres[:, 0]                          # v_t = U' diag(d_t) y~res_t
                Z[:, i * (1 + q0) + 1:(i + 1) * (1 + q0)] = d[:, None] * Q   # c_tj = U' diag(d_t) Q_t[:, j]   (:1303)
                W[i] = w
                h0.append(float(np.asarray(h0_rss).reshape(-1)[0]))
                meta.append(meta_of(t, delta, h0_rss, h0_betas))
            Zd = DeviceMatrix.from_host(ctx, Z)
            Ext = ctx.gemm(Zd, UL, ta=True)                                  # [len(ts) (1 + q0) x n] = Z' U_L
            Zd.free()
            out = ctx.emmax_scan_shared(UL, Ext, W, q0, h0, n_p)
            Ext.free()
            for i, md in enumerate(meta):
                dct = {'ps': out['ps'][i], 'f_stats': out['f_stats'][i], 'rss': out['rss'][i], 'var_perc': out['var_perc'][i]}
                dct.update(md)
                results.append(dct)
        return results

    if batch is None:
        free = ctx.device_info()['free_bytes']
        n_pad = (n + 255) // 256 * 256
        per = n * n * 8 + 7 * n_pad * n_pad + 5 * 8 * num_snps
        batch = max(1, min(T, int(0.5 * free / per)))
    for t0 in range(0, T, batch):
        ts = range(t0, min(T, t0 + batch))
        Rs, V, h0, meta = [], [], [], []
        for t in ts:
            delta = float(r['delta'][t])
            H = LazyScaledRows(UL, 1.0 / np.sqrt(eigL_vals + delta))          # :898
            mt = LinearMixedModel(Y[t], ctx=ctx)
            mt.X = X
            nf = mt._null_fit(H, project=True)                                # :1290-1303
            Rs.append(nf['R'])
            V.append(nf['Yres'].reshape(-1))
            h0.append(float(np.asarray(nf['h0_rss']).reshape(-1)[0]))
            meta.append(meta_of(t, delta, nf['h0_rss'], nf['h0_betas']))
        out = ctx.emmax_scan_multi(Rs, np.asarray(V), h0, n_p)
        for Rm in Rs:
            Rm.free()
        for i, md in enumerate(meta):
            d = {'ps': out['ps'][i], 'f_stats': out['f_stats'][i], 'rss': out['rss'][i], 'var_perc': out['var_perc'][i]}
            d.update(md)
            results.append(d)
    return results
