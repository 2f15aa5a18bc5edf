"""
ORACLE -- test infrastructure only.  NOT part of the product path.

A Python-3 / numpy-2 CPU restatement of the EMMAX hot path of bvilhjal/mixmogam
(kinship -> REML -> SNP scan, plus the permutation scan).  Every function cites
the reference file:line it follows.  Only `tests/`, `__graft_entry__.smoke()`
and `bench.py`'s cpu_baseline / `--impl reference` legs may import this module;
`mixmogam_b200/` never does.

Why a restatement: the reference is Python-2 source on top of `scipy` used as a
numpy alias (`sp.mat`, `sp.zeros`, ...), neither of which exists in this image
(Python 3.12, numpy 2.3, scipy 1.18), and it ships no tests, golden vectors or
fixtures for this path.

PARITY PIN: the reference's own sources ARE executed in the build container --
tests/golden/py2shim.py token-translates them Python 2 -> 3 in memory and
re-points `scipy` at numpy for the removed aliases -- and their outputs are
committed as tests/golden/ref_*.npz (tests/golden/make_reference_golden.py).
In dtype='single', promotion='numpy2' mode this oracle reproduces every one of
those outputs BIT FOR BIT (kinship x3, emmax, with_betas, cofactors, Z,
emma_num, snp_priors, get_REML, permutations, the hdf5_data entry points:
tests/test_reference_pin.py).  The dtype='double' mode the CUDA path is held
to is the same code with the dtype switched.

Two arithmetic modes:
  dtype='single'  reference-faithful: float32 wherever the reference hard-codes
                  'single' (linear_models.py:558,589,600,773,1283).
  dtype='double'  the same algebra in float64 -- the mode the GPU path is
                  compared against at 1e-6 relative in -log10(p).

numpy-1.x "value based" scalar promotion (what the reference ran under) is
emulated by passing Python floats (weak scalars under NEP 50) wherever the
reference mixes a float32 array with a float64 scalar.
"""
import warnings

import numpy as np
from scipy import linalg, optimize, stats


def _dt(dtype):
    if dtype in ('single', np.float32, 'float32'):
        return np.dtype(np.float32)
    if dtype in ('double', np.float64, 'float64'):
        return np.dtype(np.float64)
    return np.dtype(dtype)


# --------------------------------------------------------------------------
# kinship.py
# --------------------------------------------------------------------------

def scale_k(k, verbose=False):
    """kinship.py:94-100.  c = sum((I - 11'/n) o K); K <- (n-1)/c * K."""
    k = np.asarray(k)
    n = len(k)
    c = np.sum((np.eye(n) - (1.0 / n) * np.ones(k.shape)) * np.array(k))
    scalar = (n - 1) / c
    if verbose:
        print('Kinship scaled by: %0.4f' % scalar)
    return scalar * k


def calc_ibs_kinship(snps, snps_data_format='binary', snp_dtype='int8', dtype='single',
                     chunk_size=None, scaled=True):
    """kinship.py:14-56, literal (incl. the O(n^2) Python loop of the
    'diploid_int' branch, :33-41 -- use only at small n)."""
    num_snps = len(snps)
    num_lines = len(snps[0])
    if chunk_size is None:
        chunk_size = num_lines
    k_mat = np.zeros((num_lines, num_lines), dtype=_dt(dtype))          # :26
    for snp_i in range(0, num_snps, chunk_size):                           # :29
        snps_array = np.array(snps[snp_i:snp_i + chunk_size], dtype=snp_dtype)
        snps_array = snps_array.T                                          # :32
        if snps_data_format == 'diploid_int':
            for i in range(num_lines):
                for j in range(i):
                    bin_counts = np.bincount(np.absolute(snps_array[j] - snps_array[i]))
                    if len(bin_counts) > 1:
                        k_mat[i, j] += (bin_counts[0] + 0.5 * bin_counts[1])
                    else:
                        k_mat[i, j] += bin_counts[0]
                    k_mat[j, i] = k_mat[i, j]
        elif snps_data_format == 'binary':
            sm = snps_array * 2.0 - 1.0                                    # :43 float64
            k_mat = k_mat + sm @ sm.T                                      # :44 -> float64
        else:
            raise NotImplementedError
    if snps_data_format == 'diploid_int':
        k_mat = k_mat / float(num_snps) + np.eye(num_lines)                # :51
    elif snps_data_format == 'binary':
        k_mat = k_mat / (2 * float(num_snps)) + 0.5                        # :53
    if scaled:
        k_mat = scale_k(k_mat)
    return k_mat


def ibs_counts_diploid_vectorised(snps, chunk_size=4096):
    """Integer identity used as the CPU stand-in for kinship.py:33-41 at sizes
    where the literal double loop is infeasible:  for x in {0,1,2}
        count0 + 0.5*count1 = m - L1/2,   L1_ij = sum_s |x_si - x_sj|
    and, with thermometer planes a=[x>=1], b=[x>=2],
        L1_ij = r_i + r_j - 2*(aa' + bb')_ij,   r = rowsum(a)+rowsum(b).
    Returns the exact (half-)integer count matrix c (float64), diagonal 0
    (kinship.py:35 never fills it)."""
    snps = np.asarray(snps)
    m, n = snps.shape
    g = np.zeros((n, n), dtype=np.float64)
    r = np.zeros(n, dtype=np.float64)
    for s0 in range(0, m, chunk_size):
        x = snps[s0:s0 + chunk_size]
        a = (x >= 1).astype(np.float32)
        b = (x >= 2).astype(np.float32)
        g += (a.T @ a).astype(np.float64) + (b.T @ b).astype(np.float64)
        r += a.sum(0, dtype=np.float64) + b.sum(0, dtype=np.float64)
    l1 = r[:, None] + r[None, :] - 2.0 * g
    c = m - 0.5 * l1
    np.fill_diagonal(c, 0.0)
    return c


def calc_ibs_kinship_diploid_fast(snps, dtype='single', scaled=True):
    """kinship.py:33-41,51 through the integer identity above; reproduces the
    literal loop bit for bit while counts < 2**23 (float32 half-integers)."""
    snps = np.asarray(snps)
    m, n = snps.shape
    c = ibs_counts_diploid_vectorised(snps)
    k_mat = c.astype(_dt(dtype))
    k_mat = k_mat / float(m) + np.eye(n)
    if scaled:
        k_mat = scale_k(k_mat)
    return k_mat


def calc_ibd_kinship(snps, dtype='single', scaled=True):
    """kinship.py:59-75."""
    num_snps = len(snps)
    n_indivs = len(snps[0])
    k_mat = np.zeros((n_indivs, n_indivs), dtype=_dt(dtype))
    for i in range(0, num_snps, n_indivs):
        snps_array = np.array(snps[i:i + n_indivs])
        snps_array = snps_array.T
        norm_snps_array = (snps_array - np.mean(snps_array, 0)) / np.std(snps_array, 0)
        assert np.all(~np.isnan(norm_snps_array)), 'WTF?'                  # :67
        x = norm_snps_array.T
        k_mat += x.T @ x                                                   # :69 accumulates in dtype
    k_mat = k_mat / float(num_snps)
    if scaled:
        k_mat = scale_k(k_mat)
    return k_mat


def hdf5_ibd_kinship(chrom_snps, chrom_freqs=None, min_maf=0.1, chunk_size=1000, dtype='single'):
    """hdf5_data.py:84-115 (run_emmax) / :30-62 (calculate_ibd_kinship when
    chrom_freqs is None: no MAF filter).  chrom_snps: list of (m_c, n) arrays.
    Returns (k, n_snps)."""
    n_indivs = chrom_snps[0].shape[1]
    k_mat = np.zeros((n_indivs, n_indivs), dtype=_dt(dtype))
    n_snps = 0
    for ci, snps in enumerate(chrom_snps):
        if chrom_freqs is not None:
            freqs = np.asarray(chrom_freqs[ci])
            mafs = np.minimum(freqs, 1 - freqs)
            snps = snps[mafs > min_maf]
        num_snps = len(snps)
        for i in range(0, num_snps, chunk_size):
            end_i = min(i + chunk_size, num_snps)
            x = snps[i:end_i]
            x = x.T
            x = (x - np.mean(x, 0)) / np.std(x, 0)
            x = x.T
            n_snps += len(x)
            k_mat += np.dot(x.T, x)
    k_mat = k_mat / float(n_snps)
    c = np.sum((np.eye(len(k_mat)) - (1.0 / len(k_mat)) * np.ones(k_mat.shape)) * np.array(k_mat))
    scalar = (len(k_mat) - 1) / c
    return scalar * k_mat, n_snps


# --------------------------------------------------------------------------
# linear_models.py
# --------------------------------------------------------------------------

def _residue_value(rss):
    """The reference tests `if rss:` on what scipy.linalg.lstsq returns as
    residues (linear_models.py:1325,1329): false for an empty array (rank
    deficient column) and for an exact 0.  numpy 2 refuses the truth value of
    an empty array, so spell the test out."""
    rss = np.asarray(rss)
    if rss.size == 0:
        return None
    v = rss.reshape(-1)[0]
    if v == 0:
        return None
    return v


class LinearModel(object):
    """linear_models.py:81-130 and 196-257: the plain linear model and its SNP scan `fast_f_test` -- the EMMAX scan's
    sibling with M = I - QQ' and no kinship.  Y and X are float64 here (:90-91, unlike the mixed
    model's float32 :563-565); `dtype` is the reference's hard-coded 'single' of :202."""

    def __init__(self, Y=None, dtype='single'):
        self.dtype = _dt(dtype)
        self.n = len(Y)
        self.Y = np.array(Y, dtype=np.float64).reshape(self.n, 1)              # :90  sp.matrix(Y).T
        self.X = np.ones((self.n, 1))                                          # :91
        self.p = 1
        self.beta_est = None
        self.cofactors = []

    def add_factor(self, x, lin_depend_thres=1e-8):
        """linear_models.py:98-113."""
        new_x = np.array(x)
        new_x.shape = len(x)
        (beta, rss, rank, sigma) = linalg.lstsq(self.X, new_x)
        if float(np.asarray(rss).reshape(-1)[0] if np.asarray(rss).size else 0.0) < lin_depend_thres:
            warnings.warn('A factor was found to be linearly dependent on the factors already in the X matrix.  Hence skipping it!')
            return False
        new_x.shape = (self.n, 1)
        self.X = np.hstack([self.X, new_x])
        self.cofactors.append(x)
        self.p += 1
        return True

    def fast_f_test(self, snps, verbose=False, Z=None, with_betas=False):
        """linear_models.py:196-257 (Z is accepted and ignored there too)."""
        dtype = self.dtype                                                     # :202
        q = 1
        p = len(self.X.T) + q
        n = self.n
        n_p = n - p
        num_snps = len(snps)

        h0_X = np.array(self.X, dtype=dtype)                                   # :209
        (h0_betas, h0_rss, h0_rank, h0_s) = linalg.lstsq(h0_X, self.Y)         # :210
        Y = np.array(self.Y - h0_X @ h0_betas, dtype=dtype)                    # :211
        h0_betas = list(map(float, list(np.asarray(h0_betas).reshape(-1))))

        if not with_betas:
            (Q, R) = linalg.qr(h0_X, mode='economic')                          # :215 (qr_decomp, :68-75)
            Q = np.array(Q, dtype=dtype)
            Q2 = Q @ Q.T
            M = np.array(np.eye(n) - Q2, dtype=dtype)                          # :218
        else:
            betas_list = [h0_betas] * num_snps

        rss_list = np.repeat(h0_rss, num_snps)                                 # float64 (h0_rss is)
        chunk_size = len(Y)
        for i in range(0, len(snps), chunk_size):
            snps_chunk = np.array(snps[i:i + chunk_size])                      # :225  integer genotypes
            if with_betas:
                Xs = snps_chunk
            else:
                Xs = np.array(snps_chunk, dtype=dtype) @ M                     # :229
            for j in range(len(Xs)):
                X_j = Xs[j:j + 1]
                if with_betas:
                    (betas, rss, rk, sigma) = linalg.lstsq(np.hstack([h0_X, X_j.T]), Y)     # :232
                    if _residue_value(rss) is None:                            # :234 `if not rss: continue`
                        continue
                    betas_list[i + j] = list(map(float, list(np.asarray(betas).reshape(-1))))
                else:
                    (betas, rss, rk, sigma) = linalg.lstsq(X_j.T, Y)           # :240
                rss_list[i + j] = np.asarray(rss).reshape(-1)[0]               # :241

        rss_ratio = h0_rss / rss_list
        var_perc = 1 - 1 / rss_ratio
        f_stats = (rss_ratio - 1) * n_p / float(q)
        p_vals = stats.f.sf(f_stats, q, n_p)

        res_d = {'ps': p_vals, 'f_stats': f_stats, 'rss': rss_list, 'var_perc': var_perc,
                 'h0_rss': h0_rss, 'h0_betas': h0_betas}
        if with_betas:
            res_d['betas'] = betas_list
        return res_d


class LinearMixedModel(object):
    """linear_models.py:554-1380 (EMMAX subset)."""

    def __init__(self, Y=None, dtype='single', promotion='numpy1'):
        """linear_models.py:558-574.

        promotion: scalar type-promotion rules the float32 mode runs under.
          'numpy1' (default) -- value-based casting of numpy < 2, what the reference was written for: a
                     float64 *scalar* never upcasts a float32 array, but scalar-with-scalar arithmetic
                     (np.float32 with a Python float) gives float64, so the secant refinement of
                     delta runs in float64.
          'numpy2' -- NEP 50, what the reference's own source does when executed in this container
                     (tests/golden/py2shim.py): Python floats are weak (the secant refinement stays in
                     float32) and an np.float64 scalar upcasts a float32 array (the lls grid, :806).
                     In this mode the oracle reproduces the reference run here bit for bit
                     (tests/test_oracle_golden.py::test_oracle_matches_reference_run_*).
        With dtype='double' the two are identical."""
        assert promotion in ('numpy1', 'numpy2')
        self.promotion = promotion
        self.dtype = _dt(dtype)
        self.n = len(Y)
        self.y_var = np.var(Y, ddof=1, dtype=self.dtype)
        self.Y = np.array(Y, dtype=self.dtype).reshape(self.n, 1)
        self.X = np.ones((self.n, 1), dtype=self.dtype)
        self.p = 1
        self.beta_est = None
        self.cofactors = []
        self.random_effects = [('normal', np.identity(self.n))]

    def add_factor(self, x, lin_depend_thres=1e-8):
        """linear_models.py:98-113."""
        new_x = np.array(x)
        new_x.shape = len(x)
        (beta, rss, rank, sigma) = linalg.lstsq(self.X, new_x)
        if float(np.asarray(rss).reshape(-1)[0] if np.asarray(rss).size else 0.0) < lin_depend_thres:
            warnings.warn('A factor was found to be linearly dependent on the factors already in the X matrix.  Hence skipping it!')
            return False
        new_x.shape = (self.n, 1)
        self.X = np.hstack([self.X, new_x])
        self.cofactors.append(x)
        self.p += 1
        return True

    def add_random_effect(self, cov_matrix=None, effect_type='normal'):
        """linear_models.py:577-580 (re-scales K)."""
        if effect_type != 'normal':
            raise Exception('Currently, only Normal random effects are allowed.')
        self.random_effects.append((effect_type, scale_k(cov_matrix)))

    def _get_eigen_L_(self, K=None, dtype=None):
        """linear_models.py:589-596.  vectors are stored as ROWS (evecs.T)."""
        dtype = self.dtype if dtype is None else _dt(dtype)
        if K is None:
            K = self.random_effects[1][1]
        evals, evecs = linalg.eigh(K)
        evals = np.array(evals, dtype=dtype)
        return {'values': evals, 'vectors': np.array(evecs, dtype=dtype).T}

    def _get_eigen_R_(self, X=None, K=None, hat_matrix=None, dtype=None):
        """linear_models.py:600-615."""
        dtype = self.dtype if dtype is None else _dt(dtype)
        if X is None:
            X = self.X
        q = X.shape[1]
        if hat_matrix is None:
            X_squared_inverse = linalg.pinv(X.T @ X)
            hat_matrix = X @ X_squared_inverse @ X.T
        if K is None:
            K = self.random_effects[1][1]
        S = np.identity(self.n) - hat_matrix
        M = np.array(S @ (np.asarray(K) + self.random_effects[0][1]) @ S, dtype='double')
        evals, evecs = linalg.eigh(M, overwrite_a=True)
        eig_values = np.array(evals[q:], dtype=dtype) - 1
        return {'values': eig_values, 'vectors': (np.array(evecs, dtype=dtype).T[q:])}

    def _rell_(self, delta, eig_vals, sq_etas):
        """linear_models.py:618-623."""
        num_eig_vals = len(eig_vals)
        c_1 = 0.5 * num_eig_vals * (np.log(num_eig_vals / (2.0 * np.pi)) - 1)
        v = eig_vals + delta
        res = c_1 - 0.5 * (num_eig_vals * np.log(np.sum(sq_etas.flatten() / v)) + np.sum(np.log(v)))
        return res

    def _redll_(self, delta, eig_vals, sq_etas):
        """linear_models.py:626-631."""
        num_eig_vals = len(eig_vals)
        v1 = eig_vals + delta
        v2 = sq_etas.flatten() / v1
        res = (num_eig_vals * np.sum(v2 / v1) / np.sum(v2) - np.sum(1.0 / v1))
        return res

    def _ll_(self, delta, eig_vals, eig_vals_L, sq_etas):
        """linear_models.py:634-641."""
        n = self.n
        c_1 = 0.5 * n * (np.log(n / (2.0 * np.pi)) - 1)
        v1 = eig_vals + delta
        v2 = eig_vals_L + delta
        res = c_1 - 0.5 * (n * np.log(np.sum(sq_etas.flatten() / v1)) + np.sum(np.log(v2)))
        return res

    def _dll_(self, delta, eig_vals, eig_vals_L, sq_etas):
        """linear_models.py:644-650."""
        v1 = eig_vals + delta
        v2 = sq_etas.flatten() / v1
        v3 = eig_vals_L + delta
        res = (self.n * np.sum(v2 / v1) / np.sum(v2) - np.sum(1.0 / v3))
        return res

    def get_ML(self, ngrids=100, llim=-10, ulim=10, esp=1e-6, eig_L=None, eig_R=None, H=None, H_inv=None, H_sqrt_inv=None,
               dtype=None):
        """linear_models.py:672-696."""
        dtype = self.dtype if dtype is None else _dt(dtype)
        if H is None:
            if not eig_L:
                K = self.random_effects[1][1]
                eig_L = self._get_eigen_L_(K)
            res = self.get_estimates(eig_L, ngrids=ngrids, llim=llim, ulim=ulim, esp=esp, method='ML', eig_R=eig_R)
        else:
            evals, evecs = linalg.eigh(H)
            X_t = np.array(H_sqrt_inv @ self.X, dtype=dtype)
            Y_t = H_sqrt_inv @ self.Y
            (betas, mahalanobis_rss, rank, hs) = linalg.lstsq(X_t, Y_t)
            rss = np.sum(np.array(self.Y - np.dot(self.X, betas)) ** 2)
            n = Y_t.shape[0]
            ll = -0.5 * (n * np.log(2 * np.pi) + np.sum(np.log(evals)) + mahalanobis_rss)
            assert len(mahalanobis_rss) > 0, 'WTF?'
            res = {'ll': ll, 'rss': rss, 'mahalanobis_rss': mahalanobis_rss}
        return res

    def get_REML(self, ngrids=100, llim=-10, ulim=10, esp=1e-6, eig_L=None, eig_R=None):
        """linear_models.py:653-668."""
        if not eig_L:
            K = self.random_effects[1][1]
            eig_L = self._get_eigen_L_(K)
        res = self.get_estimates(eig_L, ngrids=ngrids, llim=llim, ulim=ulim, esp=esp, method='REML',
                                 eig_R=eig_R)
        res['eig_L'] = eig_L
        return res

    def get_estimates(self, eig_L, K=None, xs=None, ngrids=50, llim=-10, ulim=10, esp=1e-6,
                      return_pvalue=False, return_f_stat=False, method='REML', verbose=False,
                      dtype=None, eig_R=None, rss_0=None, literal_vg=None, reuse_eig_R=False):
        """linear_models.py:771-927 (REML and ML branches).

        reuse_eig_R=False is the reference: `if not (eig_R and xs != None)` at
        :787 is always true when xs is None, so a caller-supplied eig_R is
        recomputed.  reuse_eig_R=True skips that (identical values, used to
        keep big-n oracle runs affordable).
        literal_vg: evaluate the (p,1)/(p,) -> p x p broadcast of :894 literally
        (default: literal when p <= 3000, else its closed form
        sum(sq_etas)*sum(1/(lambda+delta))/p)."""
        dtype = self.dtype if dtype is None else _dt(dtype)
        if method not in ('REML', 'ML'):
            raise Exception("method must be 'REML' or 'ML'")
        if xs is not None:
            X = np.hstack([self.X, xs])
        else:
            X = self.X

        if not (eig_R and xs is not None):
            if not (reuse_eig_R and eig_R):
                eig_R = self._get_eigen_R_(X=X, K=K)
        q = X.shape[1]
        n = self.n
        p = n - q
        m = ngrids + 1

        etas = np.array(eig_R['vectors'] @ self.Y, dtype=dtype)
        sq_etas = etas * etas
        log_deltas = (np.arange(m, dtype=dtype) / ngrids) * (ulim - llim) + llim
        deltas = np.exp(log_deltas)
        assert len(deltas) == m, 'Number of deltas is incorrect.'
        eig_vals = np.array(eig_R['values'], dtype=dtype)
        assert len(eig_vals) == p, 'Number of eigenvalues is incorrect.'

        lambdas = np.reshape(np.repeat(eig_vals, m), (p, m)) + np.reshape(np.repeat(deltas, p), (m, p)).T
        s1 = np.sum(sq_etas / lambdas, axis=0)
        if method == 'REML':
            s2 = np.sum(np.log(lambdas), axis=0)
            log_c = np.log((p) / (2.0 * np.pi))                    # np.float64 scalar
            if self.promotion == 'numpy1':
                log_c = float(log_c)                               # value-based casting: does not upcast s1, s2
            lls = 0.5 * (p * (log_c - 1 - np.log(s1)) - s2)
            s3 = np.sum(sq_etas / (lambdas * lambdas), axis=0)
            s4 = np.sum(1 / lambdas, axis=0)
            dlls = 0.5 * (p * s3 / s1 - s4)
        else:                                                      # :811-824
            eig_vals_L = np.array(eig_L['values'], dtype=dtype)
            xis = np.reshape(np.repeat(eig_vals_L, m), (n, m)) + np.reshape(np.repeat(deltas, n), (m, n)).T
            s2 = np.sum(np.log(xis), axis=0)
            log_c = np.log((n) / (2.0 * np.pi))
            if self.promotion == 'numpy1':
                log_c = float(log_c)
            lls = 0.5 * (n * (log_c - 1 - np.log(s1)) - s2)
            s3 = np.sum(sq_etas / (lambdas * lambdas), axis=0)
            s4 = np.sum(1 / xis, axis=0)
            dlls = 0.5 * (n * s3 / s1 - s4)

        max_ll_i = np.argmax(lls)
        max_ll = lls[max_ll_i]

        last_dll = dlls[0]
        last_ll = lls[0]
        zero_intervals = []
        for i in range(1, len(dlls)):
            if dlls[i] < 0 and last_dll > 0:
                zero_intervals.append(((lls[i] + last_ll) * 0.5, i))
            last_ll = lls[i]
            last_dll = dlls[i]

        if len(zero_intervals) > 0:
            opt_ll, opt_i = max(zero_intervals)
            opt_delta = deltas[opt_i - 1] + deltas[opt_i]     # float32 scalar sum (:841)
            if self.promotion == 'numpy1':
                opt_delta = float(opt_delta)                   # float32 scalar * Python float -> float64 scalar
            opt_delta = 0.5 * opt_delta
            try:
                with warnings.catch_warnings():
                    warnings.simplefilter("ignore")
                    if method == 'REML':
                        new_opt_delta = optimize.newton(self._redll_, opt_delta, args=(eig_vals, sq_etas), tol=esp,
                                                        maxiter=100)
                    else:
                        new_opt_delta = optimize.newton(self._dll_, opt_delta, args=(eig_vals, eig_vals_L, sq_etas),
                                                        tol=esp, maxiter=100)
                    if self.promotion == 'numpy1':
                        new_opt_delta = float(new_opt_delta)
            except Exception:
                new_opt_delta = opt_delta
            if opt_i > 1 and deltas[opt_i - 1] - esp < new_opt_delta < deltas[opt_i] + esp:
                opt_delta = new_opt_delta
                opt_ll = self._rell_(opt_delta, eig_vals, sq_etas)
            elif opt_i == 1 and 0.0 < new_opt_delta < deltas[opt_i] + esp:
                opt_delta = new_opt_delta
                opt_ll = self._rell_(opt_delta, eig_vals, sq_etas)
            elif opt_i == len(deltas) - 1 and new_opt_delta > deltas[opt_i - 1] - esp \
                    and not np.isinf(new_opt_delta):
                opt_delta = new_opt_delta
                opt_ll = self._rell_(opt_delta, eig_vals, sq_etas)
            if method == 'REML':
                opt_ll = self._rell_(opt_delta, eig_vals, sq_etas)             # :881-882
            else:
                opt_ll = self._ll_(opt_delta, eig_vals, eig_vals_L, sq_etas)   # :883-884

            if opt_ll < max_ll:
                opt_delta = deltas[max_ll_i]
        else:
            opt_delta = deltas[max_ll_i]
            opt_ll = max_ll
        if self.promotion == 'numpy1':
            opt_delta = float(opt_delta)                       # same value; a Python float is 'weak' under NEP 50

        if literal_vg is None:
            literal_vg = p <= 3000
        if literal_vg:
            l = sq_etas / (eig_vals + opt_delta)          # :894 (p,1)/(p,) -> (p,p)  [reference bug, kept]
            opt_vg = np.sum(l) / p
        else:
            opt_vg = np.sum(sq_etas, dtype=np.float64) * np.sum(1.0 / (eig_vals + opt_delta), dtype=np.float64) / p
        opt_ve = opt_vg * opt_delta

        H_sqrt_inv = np.array(1.0 / np.sqrt(eig_L['values'] + opt_delta), dtype=dtype)[:, None] * eig_L['vectors']  # :898
        X_t = H_sqrt_inv @ X
        Y_t = H_sqrt_inv @ self.Y
        (beta_est, mahalanobis_rss, rank, sigma) = linalg.lstsq(X_t, Y_t)
        x_beta = X @ beta_est
        residuals = self.Y - x_beta
        rss = residuals.T @ residuals
        res_dict = {'max_ll': opt_ll, 'delta': opt_delta, 'beta': beta_est, 've': opt_ve, 'vg': opt_vg,
                    'rss': rss, 'mahalanobis_rss': mahalanobis_rss, 'H_sqrt_inv': H_sqrt_inv,
                    'pseudo_heritability': 1.0 / (1 + opt_delta),
                    # extras (not in the reference dict) for the parity tests
                    '_lls': lls, '_dlls': dlls, '_deltas': deltas, '_eig_R': eig_R}

        if xs is not None and return_f_stat:
            h0_X = H_sqrt_inv @ self.X
            (h0_betas, h0_rss, h0_rank, h0_s) = linalg.lstsq(h0_X, Y_t)
            f_stat = (h0_rss / mahalanobis_rss - 1) * p / xs.shape[1]
            res_dict['var_perc'] = 1.0 - mahalanobis_rss / h0_rss
            res_dict['f_stat'] = float(np.asarray(f_stat).reshape(-1)[0])
        if return_pvalue:
            p_val = stats.f.sf(f_stat, (xs.shape[1]), p)
            res_dict['p_val'] = float(np.asarray(p_val).reshape(-1)[0])
        return res_dict

    def expedited_REML_t_test(self, snps, ngrids=50, llim=-4, ulim=10, esp=1e-6, verbose=True, eig_L=None):
        """linear_models.py:931-968: exact EMMA, one eig_R per SNP."""
        K = self.random_effects[1][1]
        if eig_L is None:
            eig_L = self._get_eigen_L_(K)
        num_snps = len(snps)
        f_stats = np.empty(num_snps)
        vgs = np.empty(num_snps)
        ves = np.empty(num_snps)
        max_lls = np.empty(num_snps)
        var_perc = np.empty(num_snps)
        rss_list = np.empty(num_snps)
        betas = []
        p_vals = np.empty(num_snps)
        for i, snp in enumerate(snps):
            res = self.get_estimates(eig_L=eig_L, xs=np.array(snp).reshape(-1, 1), ngrids=ngrids, llim=llim,
                                     ulim=ulim, esp=esp, return_pvalue=True, return_f_stat=True)
            f_stats[i] = res['f_stat']
            vgs[i] = res['vg']
            ves[i] = res['ve']
            max_lls[i] = res['max_ll']
            var_perc[i] = np.asarray(res['var_perc']).reshape(-1)[0]
            betas.append(list(map(float, list(np.asarray(res['beta']).reshape(-1)))))
            p_vals[i] = res['p_val']
            rss_list[i] = np.asarray(res['rss']).reshape(-1)[0]
        return {'ps': p_vals, 'f_stats': f_stats, 'vgs': vgs, 'ves': ves, 'var_perc': var_perc,
                'max_lls': max_lls, 'betas': betas, 'rss': rss_list}

    def emmax_f_test(self, snps, snp_priors=None, Z=None, with_betas=False, method='REML',
                     eig_L=None, eig_R=None, emma_num=100):
        """linear_models.py:1233-1267."""
        if not eig_L:
            eig_L = self._get_eigen_L_()
        if not eig_R:
            eig_R = self._get_eigen_R_(X=self.X)
        res = self.get_estimates(eig_L, method=method, eig_R=eig_R, reuse_eig_R=True)
        r = self._emmax_f_test_(snps, res['H_sqrt_inv'], snp_priors=snp_priors, Z=Z, with_betas=with_betas,
                                emma_num=emma_num, eig_L=eig_L)
        r['pseudo_heritability'] = res['pseudo_heritability']
        r['ve'] = res['ve']
        r['vg'] = res['vg']
        r['max_ll'] = res['max_ll']
        r['_delta'] = res['delta']
        return r

    def _emmax_f_test_(self, snps, H_sqrt_inv, snp_priors=None, verbose=False, return_transformed_snps=False,
                       Z=None, with_betas=False, emma_num=100, eig_L=None, **kwargs):
        """linear_models.py:1272-1380."""
        dtype = self.dtype                                                  # :1283 ('single' in the reference)
        q = 1
        p = len(self.X.T) + q
        n = self.n
        n_p = n - p
        num_snps = len(snps)

        h0_X = np.array(H_sqrt_inv @ self.X, dtype=dtype)
        Y = H_sqrt_inv @ self.Y
        (h0_betas, h0_rss, h0_rank, h0_s) = linalg.lstsq(h0_X, Y)
        Y = np.array(Y - h0_X @ h0_betas, dtype=dtype)
        h0_betas = list(map(float, list(np.asarray(h0_betas).reshape(-1))))

        if Z is not None:
            H_sqrt_inv = H_sqrt_inv @ Z

        if not with_betas:
            (Q, R) = linalg.qr(h0_X, mode='economic')
            Q2 = Q @ Q.T
            M = np.array(H_sqrt_inv.T @ (np.eye(n) - Q2), dtype=dtype)
        else:
            betas_list = [h0_betas] * num_snps
            M = H_sqrt_inv.T

        rss_list = np.repeat(h0_rss, num_snps)
        if return_transformed_snps:
            t_snps = []
        if snp_priors is not None:
            snp_priors = np.array(snp_priors)
            log_h0_rss = np.log(h0_rss)
            log_bfs = np.zeros(num_snps)
        chunk_size = len(Y)
        for i in range(0, num_snps, chunk_size):
            snps_chunk = np.array(snps[i:i + chunk_size], dtype=dtype)
            Xs = snps_chunk @ M
            for j in range(len(Xs)):
                X_j = Xs[j:j + 1]
                if return_transformed_snps:
                    t_snps.append(np.array(X_j).flatten())
                if with_betas:
                    (betas, rss, rk, sigma) = linalg.lstsq(np.hstack([h0_X, X_j.T]), Y)
                    if _residue_value(rss) is not None:
                        betas_list[i + j] = list(map(float, list(np.asarray(betas).reshape(-1))))
                else:
                    (betas, rss, rk, sigma) = linalg.lstsq(X_j.T, Y)
                rv = _residue_value(rss)
                if rv is not None:
                    rss_list[i + j] = rv
                    if snp_priors is not None:
                        log_bfs[i + j] = (log_h0_rss - np.log(rss)).reshape(-1)[0]

        rss_ratio = h0_rss / rss_list
        var_perc = 1 - 1 / rss_ratio
        f_stats = (rss_ratio - 1) * n_p / float(q)
        p_vals = stats.f.sf(f_stats, q, n_p)

        res_d = {'ps': p_vals, 'f_stats': f_stats, 'rss': rss_list, 'var_perc': var_perc,
                 'h0_rss': h0_rss, 'h0_betas': h0_betas}
        if with_betas:
            res_d['betas'] = betas_list
        if return_transformed_snps:
            res_d['t_snps'] = t_snps
        if snp_priors is not None:
            bfs = np.exp((log_bfs * n - np.log(n)) * 1 / 2)
            res_d['bfs'] = bfs
            pos = bfs * snp_priors / (1 - snp_priors)
            res_d['pos'] = pos
            ppas = pos / (1 + pos)
            res_d['ppas'] = ppas

        if emma_num > 0:
            pval_indices = sorted(zip(res_d['ps'], range(len(snps))))[:emma_num]
            l = list(map(list, zip(*pval_indices)))
            top_snps = [snps[pi] for pi in l[1]]
            top_emma_res = self.expedited_REML_t_test(top_snps, eig_L=eig_L)
            for pi, pv, f, r, v in zip(l[1], top_emma_res['ps'], top_emma_res['f_stats'],
                                       top_emma_res['rss'], top_emma_res['var_perc']):
                res_d['ps'][pi] = pv
                res_d['f_stats'][pi] = f
                res_d['rss'][pi] = r
                res_d['var_perc'][pi] = v
        return res_d

    def emmax_GxT_f_test(self, snps, E, Z=None, with_betas=False, method='REML', eig_L=None, eig_R=None):
        """linear_models.py:1383-1416."""
        if not eig_L:
            eig_L = self._get_eigen_L_()
        if not eig_R:
            eig_R = self._get_eigen_R_(X=self.X)
        res = self.get_estimates(eig_L, method=method, eig_R=eig_R, reuse_eig_R=True)
        r = self._emmax_GxT_f_test_(snps, res['H_sqrt_inv'], E, Z, with_betas=with_betas, eig_L=eig_L)
        r['pseudo_heritability'] = res['pseudo_heritability']
        r['ve'] = res['ve']
        r['vg'] = res['vg']
        r['max_ll'] = res['max_ll']
        return r

    def _emmax_GxT_f_test_(self, snps, H_sqrt_inv, T, Z, verbose=False, **kwargs):
        """linear_models.py:1422-1514: per SNP the genetic model [h0_X, x~] and the full model [h0_X, x~, (x o T)~]."""
        dtype = self.dtype                                                  # :1432 ('single' in the reference)
        n = self.n
        num_snps = len(snps)
        if Z is None:
            Z = np.eye(n)
        h0_X = np.array(H_sqrt_inv @ self.X, dtype=dtype)
        Y = H_sqrt_inv @ self.Y
        (h0_betas, h0_rss, h0_rank, h0_s) = linalg.lstsq(h0_X, Y)
        Y = np.array(Y - h0_X @ h0_betas, dtype=dtype)
        h0_betas = list(map(float, list(np.asarray(h0_betas).reshape(-1))))
        T_flat = np.array(T).flatten()

        betas_g_list = [h0_betas] * num_snps
        betas_gt_list = [h0_betas] * num_snps
        M = H_sqrt_inv.T
        rss_g_list = np.repeat(h0_rss, num_snps)
        rss_gt_list = np.repeat(h0_rss, num_snps)
        chunk_size = len(Y)
        for i in range(0, num_snps, chunk_size):
            snps_chunk = np.array(snps[i:i + chunk_size], dtype=dtype) @ Z.T
            GT = np.array(np.array(snps_chunk) * T_flat) @ M
            Xs = snps_chunk @ M
            for j in range(len(Xs)):
                X_j = Xs[j:j + 1]
                GT_j = GT[j:j + 1]
                (betas_g, rss_g, p, sigma) = linalg.lstsq(np.hstack([h0_X, X_j.T]), Y)
                if _residue_value(rss_g) is not None:
                    betas_g_list[i + j] = list(map(float, list(np.asarray(betas_g).reshape(-1))))
                    rss_g_list[i + j] = np.asarray(rss_g).reshape(-1)[0]
                    (betas_gt, rss_gt, p, sigma) = linalg.lstsq(np.hstack([h0_X, X_j.T, GT_j.T]), Y)
                    if _residue_value(rss_gt) is not None:
                        betas_gt_list[i + j] = list(map(float, list(np.asarray(betas_gt).reshape(-1))))
                        rss_gt_list[i + j] = np.asarray(rss_gt).reshape(-1)[0]

        q = 1
        p = len(self.X.T) + q
        n_p = n - p
        rss_g_ratio = h0_rss / rss_g_list
        var_perc_g = 1 - 1 / rss_g_ratio
        f_stats_g = (rss_g_ratio - 1) * n_p / float(q)
        p_vals_g = stats.f.sf(f_stats_g, q, n_p)
        g_res_d = {'ps': p_vals_g, 'f_stats': f_stats_g, 'rss': rss_g_list, 'var_perc': var_perc_g,
                   'h0_rss': h0_rss, 'h0_betas': h0_betas, 'betas': betas_g_list}

        q = 2
        p = len(self.X.T) + q
        n_p = n - p
        rss_gt_ratio = h0_rss / rss_gt_list
        var_perc_gt = 1 - 1 / rss_gt_ratio
        f_stats_gt = (rss_gt_ratio - 1) * n_p / float(q)
        p_vals_gt = stats.f.sf(f_stats_gt, q, n_p)
        gt_res_d = {'ps': p_vals_gt, 'f_stats': f_stats_gt, 'rss': rss_gt_list, 'var_perc': var_perc_gt,
                    'betas': betas_gt_list}

        q = 1
        p = len(self.X.T) + q
        n_p = n - p
        rss_gt_g_ratio = rss_g_list / rss_gt_list
        var_perc_gt_g = 1 - 1 / rss_gt_g_ratio
        f_stats_gt_g = (rss_gt_g_ratio - 1) * n_p / float(q)
        p_vals_gt_g = stats.f.sf(f_stats_gt_g, q, n_p)
        gt_g_res_d = {'ps': p_vals_gt_g, 'f_stats': f_stats_gt_g, 'var_perc': var_perc_gt_g}
        return {'g_res': g_res_d, 'gt_res': gt_res_d, 'gt_g_res': gt_g_res_d}

    def _emmax_permutations_(self, snps, K, H_sqrt_inv, num_perm=100, Ys=None):
        """linear_models.py:1125-1175, quirks kept: self.Y is mean-centred in
        place (:1140), the null fit is subtracted twice (:1144,:1147), SNPs are
        centred and rotated by H' with no Q-projection (:1159-1160), Ys is
        built by cumulative in-place np.random.shuffle on the global legacy RNG
        (:1151-1154).  Pass Ys to bypass the RNG."""
        q = 1
        p = len(self.X.T) + q
        n = self.n
        n_p = n - p
        self.Y = self.Y - np.mean(self.Y)
        h0_X = np.array(H_sqrt_inv @ self.X, dtype=self.dtype)
        Y = H_sqrt_inv @ self.Y
        (h0_betas, h0_rss, h0_rank, h0_s) = linalg.lstsq(h0_X, Y)
        Y = np.array(Y - h0_X @ h0_betas, dtype=self.dtype)
        h0_betas = list(map(float, list(np.asarray(h0_betas).reshape(-1))))
        if len(h0_betas) != 1:
            raise ValueError('reference :1147 only type-checks with a single fixed effect')
        # :1147 second subtraction: h0_betas is now a Python list, `matrix * list` goes through
        # asmatrix(list) -- a float64 1x1 matrix -- so from here on Y is float64 in the reference too
        Y = Y - h0_X @ np.asarray(h0_betas, dtype=np.float64).reshape(1, 1)
        num_snps = len(snps)
        chunk_size = len(Y)
        if Ys is None:
            Ys = np.zeros((chunk_size, num_perm))
            for perm_i in range(num_perm):
                np.random.shuffle(Y)
                Ys[:, perm_i] = Y[:, 0]
        min_rss_list = np.repeat(h0_rss, num_perm)
        for i in range(0, num_snps, chunk_size):
            snps_chunk = np.array(snps[i:i + chunk_size])
            snps_chunk = snps_chunk - np.mean(snps_chunk, axis=1).reshape(-1, 1)
            Xs = snps_chunk @ (H_sqrt_inv.T)
            for j in range(len(Xs)):
                (betas, rss_list, rk, sigma) = linalg.lstsq(Xs[j:j + 1].T, Ys)
                if np.asarray(rss_list).size:
                    min_rss_list = np.minimum(min_rss_list, rss_list)
        max_f_stats = ((h0_rss / min_rss_list) - 1.0) * n_p / float(q)
        min_pvals = (stats.f.sf(max_f_stats, q, n_p))
        return {'min_ps': min_pvals, 'max_f_stats': max_f_stats, '_Ys': Ys, '_h0_rss': h0_rss}


def emmax(snps, phenotypes, K, cofactors=None, Z=None, with_betas=False, emma_num=0, dtype='single',
          promotion='numpy1'):
    """linear_models.py:1790-1816."""
    lmm = LinearMixedModel(phenotypes, dtype=dtype, promotion=promotion)
    if Z is not None:
        lmm.add_random_effect(Z @ K @ Z.T)
        if cofactors:
            for cofactor in cofactors:
                lmm.add_factor(Z @ cofactor)
    else:
        lmm.add_random_effect(K)
        if cofactors:
            for cofactor in cofactors:
                lmm.add_factor(cofactor)
    return lmm.emmax_f_test(snps, Z=Z, with_betas=with_betas, emma_num=emma_num)


# --------------------------------------------------------------------------
# Synthetic inputs shared by tests and bench
# --------------------------------------------------------------------------

def emmax_w_two_env(snps, phenotypes, K, E, cofactors=None, Z=None, dtype='single', promotion='numpy1'):
    """linear_models.py:1749-1787."""
    lmm = LinearMixedModel(phenotypes, dtype=dtype, promotion=promotion)
    if Z is not None:
        lmm.add_random_effect(Z @ K @ Z.T)
        if cofactors:
            for cofactor in cofactors:
                lmm.add_factor(Z @ cofactor)
    else:
        lmm.add_random_effect(K)
        if cofactors:
            for cofactor in cofactors:
                lmm.add_factor(cofactor)
    return lmm.emmax_GxT_f_test(snps, E=E, Z=Z)


def synth_genotypes(m, n, coding='diploid_int', seed=20240601, maf_floor=0.1):
    """SNP-major int8[m, n]; per-SNP allele frequency U(0.05, 0.5) (U(maf_floor+,0.5)
    after redraw); monomorphic / MAF<=maf_floor SNPs are redrawn."""
    rng = np.random.Generator(np.random.PCG64(seed))
    out = np.empty((m, n), dtype=np.int8)
    todo = np.arange(m)
    while todo.size:
        f = rng.uniform(0.05, 0.5, size=todo.size)
        if coding == 'diploid_int':
            x = rng.binomial(2, f[:, None], size=(todo.size, n)).astype(np.int8)
            af = x.mean(1) / 2.0
        else:
            x = (rng.random((todo.size, n)) < f[:, None]).astype(np.int8)
            af = x.mean(1)
        out[todo] = x
        maf = np.minimum(af, 1 - af)
        todo = todo[maf <= maf_floor]
    return out


def synth_phenotype(snps, K=None, n_causal=10, seed=7, h2_poly=0.5):
    """y = X'beta + g + e, standardised."""
    rng = np.random.Generator(np.random.PCG64(seed))
    m, n = snps.shape
    idx = rng.choice(m, size=min(n_causal, m), replace=False)
    beta = rng.normal(0, 0.5, size=idx.size)
    y = beta @ snps[idx].astype(np.float64)
    if K is not None:
        w, v = np.linalg.eigh(np.asarray(K, dtype=np.float64))
        g = v @ (np.sqrt(np.clip(w, 0, None)) * rng.normal(size=n))
        y = y + np.sqrt(h2_poly) * g
    y = y + rng.normal(size=n)
    y = (y - y.mean()) / y.std()
    return y
