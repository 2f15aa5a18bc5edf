"""
ctypes binding of libmixmogam_b200.so (include/mixmogam_b200.h).

There is no CPU fallback: importing this module without the built library, or creating a
context without a B200, raises.  Build the library with `python -c "import __graft_entry__ as g; g.build()"`
(or `make -C mixmogam_b200/csrc`).
"""
import ctypes as C
import os
import threading
import weakref

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libmixmogam_b200.so')

CODING_BINARY, CODING_DIPLOID = 0, 1
IMPL_AUTO, IMPL_TCGEN05, IMPL_SIMT, IMPL_DMMA = 0, 1, 2, 3
_IMPL_NAMES = {'auto': IMPL_AUTO, 'tcgen05': IMPL_TCGEN05, 'simt': IMPL_SIMT, 'dmma': IMPL_DMMA}

ERROR_NAMES = {0: 'MMG_OK', -1: 'MMG_EBADARG', -2: 'MMG_ECUDA', -3: 'MMG_ENCCL', -4: 'MMG_ECUSOLVER',
               -5: 'MMG_EOOM', -6: 'MMG_ECUBLAS', -7: 'MMG_EVALUE'}


class MmgError(RuntimeError):
    def __init__(self, code, msg):
        RuntimeError.__init__(self, '%s: %s' % (ERROR_NAMES.get(code, code), msg))
        self.code = code


# every symbol the header declares: name -> (restype, argtypes)
_c_ctx = C.c_void_p
_i64 = C.c_int64
_dp = C.POINTER(C.c_double)
_vp = C.c_void_p
SIGNATURES = {
    'mmg_create': (C.c_int, [C.c_int, C.POINTER(_c_ctx)]),
    'mmg_destroy': (C.c_int, [_c_ctx]),
    'mmg_last_error': (C.c_char_p, [_c_ctx]),
    'mmg_device_info': (C.c_int, [_c_ctx, C.c_char_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int),
                                  C.POINTER(_i64), C.POINTER(_i64)]),
    'mmg_sync': (C.c_int, [_c_ctx]),
    'mmg_stream_handle': (C.c_int, [_c_ctx, C.POINTER(_vp)]),
    'mmg_launch_count': (_i64, [_c_ctx]),
    'mmg_timer_get': (C.c_int, [_c_ctx, C.c_char_p, _dp, C.POINTER(_i64)]),
    'mmg_timer_reset': (C.c_int, [_c_ctx]),
    'mmg_last_kernel_ms': (C.c_int, [_c_ctx, C.c_char_p, _dp]),
    'mmg_last_scan_info': (C.c_int, [_c_ctx, C.POINTER(C.c_int), _dp]),
    'mmg_host_alloc': (C.c_int, [C.POINTER(_vp), _i64]),
    'mmg_host_free': (C.c_int, [_vp]),
    'mmg_mat_create': (C.c_int, [_c_ctx, _i64, _i64, C.POINTER(_i64)]),
    'mmg_mat_alloc': (C.c_int, [_c_ctx, _i64, _i64, C.c_int, C.POINTER(_i64)]),
    'mmg_mat_free': (C.c_int, [_c_ctx, _i64]),
    'mmg_mat_shape': (C.c_int, [_c_ctx, _i64, C.POINTER(_i64), C.POINTER(_i64)]),
    'mmg_mat_upload': (C.c_int, [_c_ctx, _i64, _vp, _i64]),
    'mmg_mat_download': (C.c_int, [_c_ctx, _i64, _vp, _i64]),
    'mmg_mat_download_rows': (C.c_int, [_c_ctx, _i64, _i64, _i64, _i64, _vp, _i64]),
    'mmg_mat_device_ptr': (C.c_int, [_c_ctx, _i64, C.POINTER(_vp), C.POINTER(_i64)]),
    'mmg_mat_copy': (C.c_int, [_c_ctx, _i64, _i64]),
    'mmg_mat_gemm': (C.c_int, [_c_ctx, C.c_int, C.c_int, C.c_double, _i64, _i64, C.c_double, _i64]),
    'mmg_mat_scale_rows': (C.c_int, [_c_ctx, _i64, _vp]),
    'mmg_mat_add_diag': (C.c_int, [_c_ctx, _i64, C.c_double]),
    'mmg_mat_rotation': (C.c_int, [_c_ctx, _i64, _vp, _vp, C.c_int, _i64]),
    'mmg_mat_scale_k': (C.c_int, [_c_ctx, _i64, _dp]),
    'mmg_mat_scale_k_copy': (C.c_int, [_c_ctx, _i64, _i64, _dp]),
    'mmg_mat_syevd': (C.c_int, [_c_ctx, _i64, _vp, _dp]),
    'mmg_snps_upload': (C.c_int, [_c_ctx, _vp, _i64, _i64, _i64]),
    'mmg_snps_upload_rows': (C.c_int, [_c_ctx, _vp, _i64, _i64]),
    'mmg_snps_reserve': (C.c_int, [_c_ctx, _i64, _i64]),
    'mmg_snps_write': (C.c_int, [_c_ctx, _i64, _vp, _i64, _i64]),
    'mmg_snps_free': (C.c_int, [_c_ctx]),
    'mmg_snps_shape': (C.c_int, [_c_ctx, C.POINTER(_i64), C.POINTER(_i64)]),
    'mmg_snps_device_ptr': (C.c_int, [_c_ctx, C.POINTER(_vp), C.POINTER(_i64)]),
    'mmg_snps_row_sums': (C.c_int, [_c_ctx, _vp, _vp]),
    'mmg_kinship_gram_i8': (C.c_int, [_c_ctx, C.c_int, C.c_int, _i64, _i64, C.c_int]),
    'mmg_kinship_gram_i8_host': (C.c_int, [_c_ctx, C.c_int, C.c_int, C.c_void_p, _i64, _i64, _i64, C.c_int]),
    'mmg_host_pack2': (C.c_int, [C.c_void_p, _i64, _i64, _i64, C.c_void_p, _i64, C.c_int]),
    'mmg_kinship_gram_i8_host_packed2': (C.c_int, [_c_ctx, C.c_int, C.c_int, C.c_void_p, _i64, _i64, _i64, C.c_int]),
    'mmg_snps_upload_packed2': (C.c_int, [_c_ctx, C.c_void_p, _i64, _i64, _i64]),
    'mmg_host_threads_default': (C.c_int, []),
    'mmg_last_h2d_info': (C.c_int, [_c_ctx, C.POINTER(_i64), C.POINTER(_i64), _dp]),
    'mmg_kinship_gram_ptr': (C.c_int, [_c_ctx, C.POINTER(_vp), C.POINTER(_i64), C.POINTER(_i64)]),
    'mmg_kinship_gram_download': (C.c_int, [_c_ctx, _vp]),
    'mmg_kinship_gram_tri': (C.c_int, [_c_ctx, C.c_int, C.POINTER(_vp), C.POINTER(_i64)]),
    'mmg_kinship_finalize_f64': (C.c_int, [_c_ctx, C.c_int, _i64, C.c_int, _i64, _dp]),
    'mmg_kinship_ibd_accumulate_f64': (C.c_int, [_c_ctx, _i64, _i64, _i64, _vp, C.POINTER(_i64)]),
    'mmg_reml_f64': (C.c_int, [_c_ctx, _vp, _vp, _i64, _i64, _vp, _i64, C.c_double, _vp, _vp, _vp, _vp, _vp]),
    'mmg_emma_f64': (C.c_int, [_c_ctx, C.c_int, _i64, _vp, _vp, C.c_int, _vp, _vp, _vp, _i64, _vp, C.c_int, C.c_double, _vp, _vp, _vp]),
    'mmg_emmax_scan_f64': (C.c_int, [_c_ctx, _i64, _vp, C.c_int, C.c_double, C.c_double, C.c_int, _i64, _i64,
                                     _vp, _vp, _vp, _vp, _vp, _vp]),
    'mmg_emmax_scan_betas_f64': (C.c_int, [_c_ctx, _i64, _vp, C.c_int, _vp, _vp, C.c_double, _vp, C.c_double, C.c_double, C.c_int, _i64, _i64,
                                           _vp, _vp, _vp, _vp, _vp]),
    'mmg_emmax_scan_rows_f64': (C.c_int, [_c_ctx, _i64, _vp, C.c_int, C.c_double, C.c_double, _vp, _i64, _i64,
                                          _vp, _vp, _vp, _vp, _vp, _vp]),
    'mmg_emmax_scan_quad_f64': (C.c_int, [_c_ctx, _i64, _vp, C.c_double, C.c_double, _i64, _i64, _vp, _vp, _vp, _vp, _vp]),
    'mmg_scan_prepass_begin': (C.c_int, [_c_ctx, _i64, _vp, _i64, _i64]),
    'mmg_quad_form_slots': (_i64, [_i64]),
    'mmg_quad_form_tiles': (C.c_int, [_c_ctx, _i64, _i64, _i64, _i64, _dp]),
    'mmg_emmax_scan_quad_dev': (C.c_int, [_c_ctx, _i64, C.c_int, C.c_double, _i64, C.c_double, C.c_double, _i64, _i64, _i64]),
    'mmg_emmax_scan_multi_f64': (C.c_int, [_c_ctx, _vp, C.c_int, _vp, _vp, C.c_double, _i64, _i64, _vp, _vp, _vp, _vp, _vp]),
    'mmg_emmax_scan_shared_f64': (C.c_int, [_c_ctx, _i64, _i64, _vp, C.c_int, C.c_int, _vp, C.c_double, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp]),
    'mmg_emmax_perm_scan_f64': (C.c_int, [_c_ctx, _i64, _i64, C.c_int, C.c_int, _i64, _i64, _vp]),
    'mmg_f_sf_f64': (C.c_int, [_c_ctx, _vp, _i64, C.c_double, C.c_double, _vp]),
}
# libmixmogam_b200_bench.so (include/mixmogam_b200_bench.h): diagnostics, loaded on first use only
BENCH_LIB_PATH = os.path.join(_HERE, 'libmixmogam_b200_bench.so')
BENCH_SIGNATURES = {
    'mmg_microbench': (C.c_int, [_c_ctx, C.c_char_p, _dp]),
}

_lib = None
_bench_lib = None
_lib_lock = threading.Lock()


def load_bench_library():
    """dlopen the microbenchmark library (bench.py's roofline denominators); the product never needs it."""
    global _bench_lib
    load_library()
    with _lib_lock:
        if _bench_lib is None:
            if not os.path.exists(BENCH_LIB_PATH):
                raise ImportError('mixmogam_b200: %s is missing -- build it first (__graft_entry__.build())' % BENCH_LIB_PATH)
            lib = C.CDLL(BENCH_LIB_PATH)
            for name, (res, args) in BENCH_SIGNATURES.items():
                fn = getattr(lib, name)
                fn.restype = res
                fn.argtypes = args
            _bench_lib = lib
    return _bench_lib


def load_library():
    """dlopen the shared library and attach the prototypes (no GPU needed for this)."""
    global _lib
    with _lib_lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise ImportError('mixmogam_b200: %s is missing -- build it first (__graft_entry__.build()); '
                                  'there is no CPU fallback' % LIB_PATH)
            lib = C.CDLL(LIB_PATH)
            for name, (res, args) in SIGNATURES.items():
                fn = getattr(lib, name)
                fn.restype = res
                fn.argtypes = args
            _lib = lib
    return _lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def impl_id(impl):
    if isinstance(impl, str):
        return _IMPL_NAMES[impl.lower()]
    return int(impl)


class DeviceMatrix(object):
    """FP64 row-major matrix resident in HBM (mmg_mat handle)."""

    def __init__(self, ctx, rows, cols, zero=True):
        self.ctx = ctx
        self.shape = (int(rows), int(cols))
        h = _i64(0)
        ctx._ck(ctx.lib.mmg_mat_alloc(ctx.h, rows, cols, int(bool(zero)), C.byref(h)))
        self.handle = h.value

    @classmethod
    def from_host(cls, ctx, a):
        a = np.ascontiguousarray(np.asarray(a, dtype=np.float64))
        if a.ndim == 1:
            a = a.reshape(-1, 1)
        m = cls(ctx, a.shape[0], a.shape[1])
        m.upload(a)
        return m

    def upload(self, a):
        a = np.ascontiguousarray(np.asarray(a, dtype=np.float64))
        assert a.shape == self.shape, (a.shape, self.shape)
        self.ctx._ck(self.ctx.lib.mmg_mat_upload(self.ctx.h, self.handle, _ptr(a), a.shape[1]))

    def download(self, out=None, pinned=False):
        if out is None:
            out = pinned_empty(self.shape, np.float64) if pinned else np.empty(self.shape, dtype=np.float64)
        self.ctx._ck(self.ctx.lib.mmg_mat_download(self.ctx.h, self.handle, _ptr(out), out.shape[1]))
        return out

    def copy(self):
        m = DeviceMatrix(self.ctx, *self.shape)
        self.ctx._ck(self.ctx.lib.mmg_mat_copy(self.ctx.h, m.handle, self.handle))
        return m

    def download_rows(self, row0, row_step, nrows, out=None):
        """Rows row0, row0 + row_step, ... (nrows of them) as an [nrows x cols] host array (one strided copy)."""
        if out is None:
            out = result_empty((int(nrows), self.shape[1]))
        self.ctx._ck(self.ctx.lib.mmg_mat_download_rows(self.ctx.h, self.handle, int(row0), int(row_step), int(nrows), _ptr(out),
                                                        out.shape[1]))
        return out

    def device_ptr(self, sync=True):
        """(device pointer, leading dimension).  sync=False: the caller orders its work on the context's stream
        (parallel.on_lib_stream) instead of waiting for the stream to drain."""
        p, ld = C.c_void_p(0), _i64(0)
        self.ctx._ck(self.ctx.lib.mmg_mat_device_ptr(self.ctx.h, self.handle, C.byref(p), C.byref(ld)))
        if sync:
            self.ctx.sync()
        return p.value, ld.value

    def free(self):
        if self.handle and self.ctx.h:
            self.ctx.lib.mmg_mat_free(self.ctx.h, self.handle)
        self.handle = 0

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class LazyHostArray(object):
    """A device matrix that turns into a numpy array on demand (np.asarray(x), x[...], x.shape ...).
    Returned wherever the reference returns an n x n matrix the caller rarely touches
    (H_sqrt_inv, eigenvectors, the scaled kinship), so the 0.8 GB download only happens when asked for."""

    def __init__(self, dev, rows=None):
        self.dev = dev
        self._rows = rows      # optional row slice (start, stop) applied on materialisation
        self._host = None

    @property
    def shape(self):
        if self._rows is None:
            return self.dev.shape
        return (self._rows[1] - self._rows[0], self.dev.shape[1])

    def host(self):
        if self._host is None:
            a = self.dev.download()
            if self._rows is not None:
                a = a[self._rows[0]:self._rows[1]]
            self._host = a
        return self._host

    def __array__(self, dtype=None, copy=None):
        a = self.host()
        return a if dtype is None else a.astype(dtype)

    def __getitem__(self, k):
        return self.host()[k]

    def __len__(self):
        return self.shape[0]

    @property
    def T(self):
        return self.host().T

    def __matmul__(self, o):
        return self.host() @ np.asarray(o)

    def __rmatmul__(self, o):
        return np.asarray(o) @ self.host()


class LazyScaledRows(LazyHostArray):
    """H_sqrt_inv = diag(d) U (linear_models.py:898) kept as its two factors: U stays the eigenbasis already in HBM, d is a
    host vector.  The n x n product is formed only if somebody asks for it (np.asarray(x), x.dev); the scan builds its
    rotation (I - QQ') diag(d) U straight from the factors in one pass (Context.rotation)."""

    def __init__(self, U, d):
        self.U = U
        self.d = np.ascontiguousarray(d, dtype=np.float64)
        self._rows = None
        self._host = None
        self._dev = None

    @property
    def dev(self):
        if self._dev is None:
            self._dev = self.U.ctx.rotation(self.U, self.d)
        return self._dev

    @property
    def shape(self):
        return self.U.shape

    def times(self, B):
        """diag(d) U B for a host matrix B [n x c] (c small): one skinny GEMM on the device, rows scaled on the host.  The
        last product is remembered: get_estimates and the scan set-up both ask for H [X, Y] (linear_models.py:899-900, :1290)."""
        B = np.ascontiguousarray(B, dtype=np.float64)
        last = getattr(self, '_last_times', None)
        if last is not None and last[0].shape == B.shape and np.array_equal(last[0], B):
            return last[1].copy()
        Bd = DeviceMatrix.from_host(self.U.ctx, B)
        t = self.U.ctx.gemm(self.U, Bd).download() * self.d[:, None]
        Bd.free()
        self._last_times = (B.copy(), t)
        return t.copy()


class Context(object):
    """One per GPU (mmg_ctx).  Holds the resident genotype block and the device matrices."""

    def __init__(self, device=0):
        self.lib = load_library()
        self.h = None
        h = _c_ctx(0)
        rc = self.lib.mmg_create(int(device), C.byref(h))
        if rc != 0:
            raise MmgError(rc, self.lib.mmg_last_error(None).decode())
        self.h = h
        self.device = int(device)
        self._snps_key = None
        self._snps_owner = None
        self._resident = {}

    def _ck(self, rc):
        if rc != 0:
            raise MmgError(rc, self.lib.mmg_last_error(self.h).decode())

    def close(self):
        if self.h:
            self.lib.mmg_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- info / timers ----
    def device_info(self):
        name = C.create_string_buffer(64)
        sm, maj, mnr = C.c_int(0), C.c_int(0), C.c_int(0)
        fr, tot = _i64(0), _i64(0)
        self._ck(self.lib.mmg_device_info(self.h, name, C.byref(sm), C.byref(maj), C.byref(mnr), C.byref(fr), C.byref(tot)))
        return {'name': name.value.decode(), 'sm_count': sm.value, 'cc': (maj.value, mnr.value),
                'free_bytes': fr.value, 'total_bytes': tot.value}

    def launch_count(self):
        return int(self.lib.mmg_launch_count(self.h))

    def timer(self, name):
        s, c = C.c_double(0), _i64(0)
        self._ck(self.lib.mmg_timer_get(self.h, name.encode(), C.byref(s), C.byref(c)))
        return s.value, c.value

    STAGES = ('h2d', 'pack', 'gram', 'finalize', 'ibd', 'syevd', 'reml', 'matrix', 'scan_prep', 'scan', 'd2h')
    DETAIL = ('qf_gemm', 'host_qf_alloc', 'host_qf_total', 'host_qf_tiles')      # inside 'scan_prep' / host wall clock: not additive

    def timers(self, detail=False):
        return {k: self.timer(k)[0] for k in (self.STAGES + (self.DETAIL if detail else ()))}

    def timer_reset(self):
        self._ck(self.lib.mmg_timer_reset(self.h))

    def last_kernel_ms(self, which):
        v = C.c_double(0)
        self._ck(self.lib.mmg_last_kernel_ms(self.h, which.encode(), C.byref(v)))
        return v.value

    def last_scan_info(self):
        """(digit planes used, certified max relative truncation bound on x~.x~) of the most recent int8 scan."""
        k, rho = C.c_int(0), C.c_double(0)
        self._ck(self.lib.mmg_last_scan_info(self.h, C.byref(k), C.byref(rho)))
        return k.value, rho.value

    def last_h2d_info(self):
        """(chunks sent packed, chunks sent unpacked, host packing rate in GB/s) of the most recent streamed Gram."""
        p, r, g = _i64(0), _i64(0), C.c_double(0)
        self._ck(self.lib.mmg_last_h2d_info(self.h, C.byref(p), C.byref(r), C.byref(g)))
        return p.value, r.value, g.value

    def microbench(self, which):
        """Pipe-rate microbenchmarks of libmixmogam_b200_bench.so (see include/mixmogam_b200_bench.h)."""
        v = C.c_double(0)
        self._ck(load_bench_library().mmg_microbench(self.h, which.encode(), C.byref(v)))
        return v.value

    def sync(self):
        self._ck(self.lib.mmg_sync(self.h))

    def stream_ptr(self):
        """cudaStream_t of this context as an integer (mmg_stream_handle)."""
        p = C.c_void_p(0)
        self._ck(self.lib.mmg_stream_handle(self.h, C.byref(p)))
        return int(p.value or 0)

    # ---- genotypes ----
    # Residency contract: the device copy of a genotype array is reused by later calls ONLY when the array is read-only
    # (arr.flags.writeable == False; mixmogam_b200.resident(arr) marks it) -- a read-only buffer cannot have been edited in
    # place between two calls, so (address, shape) identifies its contents.  A writable array is uploaded again by every
    # call that needs it: an in-place edit (imputation, allele flip, a reused buffer) is always seen.
    def invalidate_snps(self):
        self._snps_key = None

    def _array_key(self, snps):
        a = snps
        if a.dtype != np.int8:
            a = _as_int8(a)
        if not a.flags.c_contiguous:
            a = np.ascontiguousarray(a)
        key = ('arr', a.ctypes.data, a.shape) if (a is snps and not snps.flags.writeable) else None
        return a, key

    def ensure_snps(self, snps):
        """Make `snps` (list of m int8 rows or an (m, n) array, SNP-major; kinship.py:21-23) the resident
        genotype block.  Uploads it unless this very read-only buffer is already resident (see the residency
        contract above).  Returns (m, n)."""
        if isinstance(snps, ResidentSnps):
            if snps.shape != self.snps_shape():
                raise ValueError('the resident genotype block is %r, not the %r this handle was made for' % (self.snps_shape(), snps.shape))
            return snps.shape
        if isinstance(snps, PackedGenotypes):
            key = snps._key()
            if key is None or key != self._snps_key:
                self._snps_key = None
                self._ck(self.lib.mmg_snps_upload_packed2(self.h, _ptr(snps.packed), snps.shape[0], snps.shape[1], snps.packed.shape[1]))
                self._snps_key = key
                self._snps_owner = snps if key is not None else None
            return snps.shape
        if isinstance(snps, np.ndarray) and snps.ndim == 2:
            a, key = self._array_key(snps)
            if key is None or key != self._snps_key:
                self._snps_key = None
                self._ck(self.lib.mmg_snps_upload(self.h, _ptr(a), a.shape[0], a.shape[1], a.shape[1]))
                self._snps_key = key
                self._snps_owner = snps if key is not None else None      # keeps the buffer (hence its address) alive
            return a.shape
        m = len(snps)
        if m == 0:
            raise ValueError('no SNPs')
        n = len(snps[0])
        rows_ok = all(isinstance(r, np.ndarray) and r.dtype == np.int8 and r.ndim == 1 and r.flags.c_contiguous
                      and r.shape[0] == n for r in snps)
        if not rows_ok:
            return self.ensure_snps(_as_int8(np.asarray(snps)))
        ptrs = np.fromiter((r.ctypes.data for r in snps), dtype=np.uint64, count=m)
        frozen = not any(r.flags.writeable for r in snps)
        key = ('rows', m, n, hash(ptrs.tobytes())) if frozen else None
        if key is None or key != self._snps_key:
            self._snps_key = None
            self._ck(self.lib.mmg_snps_upload_rows(self.h, _ptr(ptrs), m, n))
            self._snps_key = key
            self._snps_owner = list(snps) if key is not None else None
        return (m, n)

    def snps_reserve(self, m, n):
        """Allocate (or reuse) the resident genotype block for m SNPs x n individuals; its contents are undefined until written."""
        self._snps_key = None
        self._ck(self.lib.mmg_snps_reserve(self.h, int(m), int(n)))

    def snps_write(self, row0, rows):
        """Copy int8 rows [k x n] (C-contiguous; page-locked for the full PCIe rate) into resident rows [row0, row0 + k)."""
        rows = np.ascontiguousarray(rows, dtype=np.int8)
        self._snps_key = None
        self._ck(self.lib.mmg_snps_write(self.h, int(row0), _ptr(rows), rows.shape[0], rows.shape[1]))

    def snps_shape(self):
        m, n = _i64(0), _i64(0)
        self._ck(self.lib.mmg_snps_shape(self.h, C.byref(m), C.byref(n)))
        return m.value, n.value

    def snps_row_sums(self, with_sumsq=False):
        m, n = self.snps_shape()
        s = np.empty(m, dtype=np.int64)
        q = np.empty(m, dtype=np.int64) if with_sumsq else None
        self._ck(self.lib.mmg_snps_row_sums(self.h, _ptr(s), _ptr(q)))
        return (s, q) if with_sumsq else s

    # ---- matrices ----
    def matrix(self, rows, cols):
        return DeviceMatrix(self, rows, cols)

    def to_device(self, a):
        if isinstance(a, LazyHostArray):
            if a._rows is None:
                return a.dev
            a = a.host()
        if isinstance(a, DeviceMatrix):
            return a
        hit = self.lookup_resident(a)
        if hit is not None:
            return hit.copy()
        return DeviceMatrix.from_host(self, a)

    # ---- host arrays this context produced and still holds on the device (the kinship handed back by
    #      calc_ibs_kinship is usually passed straight into add_random_effect: skip the 0.8 GB round trip) ----
    def remember_resident(self, host, dev):
        host.flags.writeable = False
        base = np.asarray(host)
        key = (base.ctypes.data, base.shape)
        self._resident[key] = (dev, weakref.ref(base.base if base.base is not None else base))
        while len(self._resident) > 2:
            k0 = next(iter(self._resident))
            self._resident.pop(k0)[0].free()

    def lookup_resident(self, a):
        # only read-only arrays qualify: the host copy cannot have diverged from the device copy
        if (not self._resident or not isinstance(a, np.ndarray) or a.dtype != np.float64 or not a.flags.c_contiguous
                or a.flags.writeable):
            return None
        base = np.asarray(a)
        ent = self._resident.get((base.ctypes.data, base.shape))
        if ent is None:
            return None
        dev, ref = ent
        if ref() is None or not dev.handle:
            self._resident.pop((base.ctypes.data, base.shape), None)
            dev.free()
            return None
        return dev

    def gemm(self, A, B, C_out=None, ta=False, tb=False, alpha=1.0, beta=0.0):
        m = A.shape[1] if ta else A.shape[0]
        n = B.shape[0] if tb else B.shape[1]
        if C_out is None:
            C_out = DeviceMatrix(self, m, n)
        self._ck(self.lib.mmg_mat_gemm(self.h, int(ta), int(tb), alpha, A.handle, B.handle, beta, C_out.handle))
        return C_out

    def scan_prepass_begin(self, R, yres, snp_begin=0, snp_count=None):
        """Starts the scan's linear pre-pass for these resident rows on the side stream (mmg_scan_prepass_begin)."""
        m, n = self.snps_shape()
        if snp_count is None:
            snp_count = m - snp_begin
        yres = np.ascontiguousarray(np.asarray(yres, dtype=np.float64).reshape(-1))
        assert yres.shape[0] == R.shape[0]
        self._ck(self.lib.mmg_scan_prepass_begin(self.h, R.handle, _ptr(yres), int(snp_begin), int(snp_count)))

    def quad_form_slots(self, n):
        """Number of 256 x 256 blocks in the packed lower triangle of an n x n quadratic form (mmg_quad_form_slots)."""
        return int(self.lib.mmg_quad_form_slots(int(n)))

    def quad_form_tiles(self, R, slot_begin, slot_count, A_packed):
        """Blocks [slot_begin, +slot_count) of A = R'R (int8 digit-plane products) into the packed matrix A_packed
        ([>= slots x 65536]); returns the rigorous absolute error bound of the entries."""
        err = C.c_double(0)
        self._ck(self.lib.mmg_quad_form_tiles(self.h, R.handle, int(slot_begin), int(slot_count), A_packed.handle, C.byref(err)))
        return err.value

    def scale_rows(self, A, d):
        d = np.ascontiguousarray(d, dtype=np.float64)
        assert d.shape[0] == A.shape[0]
        self._ck(self.lib.mmg_mat_scale_rows(self.h, A.handle, _ptr(d)))

    def rotation(self, U, d, Q=None):
        """R = (I - QQ') diag(d) U as a new DeviceMatrix (Q None: diag(d) U), one pass over U (mmg_mat_rotation)."""
        d = np.ascontiguousarray(d, dtype=np.float64)
        q = 0
        if Q is not None:
            Q = np.ascontiguousarray(Q, dtype=np.float64).reshape(d.shape[0], -1)
            q = Q.shape[1]
        R = DeviceMatrix(self, U.shape[0], U.shape[1], zero=False)
        self._ck(self.lib.mmg_mat_rotation(self.h, U.handle, _ptr(d), _ptr(Q), q, R.handle))
        return R

    def add_diag(self, A, alpha):
        self._ck(self.lib.mmg_mat_add_diag(self.h, A.handle, float(alpha)))

    def scale_k(self, K):
        s = C.c_double(0)
        self._ck(self.lib.mmg_mat_scale_k(self.h, K.handle, C.byref(s)))
        return s.value

    def scale_k_copy(self, K):
        """scale_k(K) as a new DeviceMatrix; K is left as it is.  Stream ordered (nothing waits for the factor)."""
        out = DeviceMatrix(self, K.shape[0], K.shape[1], zero=False)
        self._ck(self.lib.mmg_mat_scale_k_copy(self.h, K.handle, out.handle, None))
        return out

    def syevd(self, A):
        """In place: A <- eigenvectors as rows; returns ascending eigenvalues."""
        w = np.empty(A.shape[0], dtype=np.float64)
        self._ck(self.lib.mmg_mat_syevd(self.h, A.handle, _ptr(w), None))
        return w

    # ---- stage 1 ----
    def kinship_gram(self, coding, impl=IMPL_AUTO, snp_begin=0, snp_count=None, reset=True):
        m, n = self.snps_shape()
        if snp_count is None:
            snp_count = m - snp_begin
        self._ck(self.lib.mmg_kinship_gram_i8(self.h, coding, impl_id(impl), snp_begin, snp_count, int(bool(reset))))

    def kinship_gram_from(self, snps, coding, impl=IMPL_AUTO):
        """Integer Gram of all of `snps`, which become the resident genotype block.  A 2-D array that is not resident yet
        is streamed: its 65 536-SNP chunks are copied on a second stream while the Gram of the chunks that have landed
        runs (mmg_kinship_gram_i8_host); anything else is uploaded first (ensure_snps).  Returns (m, n)."""
        if isinstance(snps, PackedGenotypes):
            key = snps._key()
            if key is None or key != self._snps_key:
                self._snps_key = None
                self._ck(self.lib.mmg_kinship_gram_i8_host_packed2(self.h, coding, impl_id(impl), _ptr(snps.packed), snps.shape[0], snps.shape[1],
                                                                   snps.packed.shape[1], 1))
                self._snps_key = key
                self._snps_owner = snps if key is not None else None
                return snps.shape
        if isinstance(snps, np.ndarray) and snps.ndim == 2 and snps.shape[0] > 0:
            a, key = self._array_key(snps)
            if key is None or key != self._snps_key:
                self._snps_key = None
                self._ck(self.lib.mmg_kinship_gram_i8_host(self.h, coding, impl_id(impl), _ptr(a), a.shape[0], a.shape[1],
                                                           a.shape[1], 1))
                self._snps_key = key
                self._snps_owner = snps if key is not None else None
                return a.shape
        shape = self.ensure_snps(snps)
        self.kinship_gram(coding, impl=impl, reset=True)
        return shape

    def kinship_gram_download(self):
        m, n = self.snps_shape()
        g = np.empty((n, n), dtype=np.int32)
        self._ck(self.lib.mmg_kinship_gram_download(self.h, _ptr(g)))
        return g

    def kinship_gram_ptr(self):
        p, n, ld = C.c_void_p(0), _i64(0), _i64(0)
        self._ck(self.lib.mmg_kinship_gram_ptr(self.h, C.byref(p), C.byref(n), C.byref(ld)))
        return p.value, n.value, ld.value

    def kinship_gram_tri(self, direction):
        """direction 0: pack the valid blocks of the Gram into one contiguous int32 buffer -> (device pointer, element count);
        direction 1: unpack that buffer into the Gram again (mmg_kinship_gram_tri)."""
        p, cnt = C.c_void_p(0), _i64(0)
        self._ck(self.lib.mmg_kinship_gram_tri(self.h, int(direction), C.byref(p), C.byref(cnt)))
        return p.value, cnt.value

    def kinship_finalize(self, coding, m_total, scaled, K=None):
        m, n = self.snps_shape()
        if K is None:
            K = DeviceMatrix(self, n, n)
        s = C.c_double(1.0)
        self._ck(self.lib.mmg_kinship_finalize_f64(self.h, coding, int(m_total), int(bool(scaled)), K.handle, C.byref(s)))
        return K, s.value

    def kinship_ibd_accumulate(self, K, snp_begin, snp_count, mask=None):
        used = _i64(0)
        if mask is not None:
            mask = np.ascontiguousarray(mask, dtype=np.uint8)
            assert mask.shape[0] == snp_count
        self._ck(self.lib.mmg_kinship_ibd_accumulate_f64(self.h, K.handle, snp_begin, snp_count, _ptr(mask), C.byref(used)))
        return used.value

    # ---- stage 2 ----
    def reml(self, eig_vals, sq_etas, deltas, esp):
        eig_vals = np.ascontiguousarray(eig_vals, dtype=np.float64)
        sq_etas = np.ascontiguousarray(sq_etas, dtype=np.float64)
        if sq_etas.ndim == 1:
            sq_etas = sq_etas.reshape(1, -1)
        T, p = sq_etas.shape
        assert eig_vals.shape[0] == p
        deltas = np.ascontiguousarray(deltas, dtype=np.float64)
        g = deltas.shape[0]
        lls = np.empty((T, g))
        dlls = np.empty((T, g))
        od = np.empty(T)
        ol = np.empty(T)
        fl = np.empty(T, dtype=np.int32)
        self._ck(self.lib.mmg_reml_f64(self.h, _ptr(eig_vals), _ptr(sq_etas), p, T, _ptr(deltas), g, float(esp),
                                       _ptr(lls), _ptr(dlls), _ptr(od), _ptr(ol), _ptr(fl)))
        return {'lls': lls, 'dlls': dlls, 'delta': od, 'll': ol, 'flags': fl}

    EMMA_KEYS = ('delta', 'max_ll', 'vg', 've', 'f_stat', 'p_val', 'var_perc', 'rss', 'mahalanobis_rss')

    def emma(self, UL, lam, X0, y, xs=None, snp_rows=None, deltas=None, esp=1e-6, method='REML', want_grid=False):
        """Variance-component fit(s) in the eigenbasis of K alone (mmg_emma_f64): for every SNP of `xs` ([k x n]) or every
        resident row of `snp_rows`, or -- with neither -- for the model without a SNP.  Returns a dict of length-k arrays
        (EMMA_KEYS) plus 'betas' [k x q] (and 'lls', 'dlls' [k x g] with want_grid)."""
        lam = np.ascontiguousarray(lam, dtype=np.float64)
        n = lam.shape[0]
        X0 = np.ascontiguousarray(np.asarray(X0, dtype=np.float64).reshape(n, -1))
        q0 = X0.shape[1]
        y = np.ascontiguousarray(np.asarray(y, dtype=np.float64).reshape(n))
        deltas = np.ascontiguousarray(deltas, dtype=np.float64)
        k = 0
        rows = None
        if xs is not None:
            xs = np.ascontiguousarray(np.asarray(xs, dtype=np.float64).reshape(-1, n))
            k = xs.shape[0]
        elif snp_rows is not None:
            rows = np.ascontiguousarray(snp_rows, dtype=np.int64)
            k = rows.shape[0]
        kk = max(k, 1)
        q = q0 + (1 if k else 0)
        out = np.empty((kk, len(self.EMMA_KEYS) + q))
        g = deltas.shape[0]
        lls = np.empty((kk, g)) if want_grid else None
        dlls = np.empty((kk, g)) if want_grid else None
        self._ck(self.lib.mmg_emma_f64(self.h, 1 if method == 'ML' else 0, UL.handle, _ptr(lam), _ptr(X0), q0, _ptr(y), _ptr(xs), _ptr(rows),
                                       k, _ptr(deltas), g, float(esp), _ptr(out), _ptr(lls), _ptr(dlls)))
        res = {name: out[:, i].copy() for i, name in enumerate(self.EMMA_KEYS)}
        res['betas'] = out[:, len(self.EMMA_KEYS):].copy()
        if want_grid:
            res['lls'], res['dlls'] = lls, dlls
        return res

    # ---- stage 3 ----
    def emmax_scan(self, R, V, h0_rss, n_p, impl=IMPL_AUTO, snp_begin=0, snp_count=None, want_dots=False,
                   want_stats=True):
        m, n = self.snps_shape()
        if snp_count is None:
            snp_count = m - snp_begin
        V = np.ascontiguousarray(np.asarray(V, dtype=np.float64))
        if V.ndim == 1:
            V = V.reshape(1, -1)
        nv = V.shape[0]
        assert V.shape[1] == R.shape[0], (V.shape, R.shape)
        # one page-locked slab for the per-SNP vectors (a single pooled allocation per call)
        keys = (('ps', 'f_stats', 'rss', 'var_perc') if want_stats else ()) + ('xx',)
        slab = result_empty((len(keys), snp_count))
        out = {k: slab[i] for i, k in enumerate(keys)}
        if want_dots:
            out['dots'] = result_empty((snp_count, nv))
        self._ck(self.lib.mmg_emmax_scan_f64(self.h, R.handle, _ptr(V), nv, float(h0_rss), float(n_p), impl_id(impl),
                                             snp_begin, snp_count, _ptr(out.get('ps')), _ptr(out.get('f_stats')),
                                             _ptr(out.get('rss')), _ptr(out.get('var_perc')), _ptr(out['xx']),
                                             _ptr(out.get('dots'))))
        return out

    def emmax_scan_betas(self, R, Yres, h0_X, h0_betas, h0_rss, n_p, impl=IMPL_AUTO, snp_begin=0, snp_count=None):
        """with_betas=True (linear_models.py:1323): lstsq([h0_X, x~], y~res) for every resident SNP, finished on the device
        (mmg_emmax_scan_betas_f64).  Returns ps, f_stats, rss, var_perc and betas [snp_count x (q0 + 1)] (last entry NaN where
        the SNP kept the null fit)."""
        m, n = self.snps_shape()
        if snp_count is None:
            snp_count = m - snp_begin
        h0_X = np.ascontiguousarray(h0_X, dtype=np.float64)
        Yres = np.ascontiguousarray(Yres, dtype=np.float64).reshape(-1)
        q0 = h0_X.shape[1]
        V = np.ascontiguousarray(np.vstack([Yres[None, :], h0_X.T]))
        Ainv = np.ascontiguousarray(np.linalg.inv(h0_X.T @ h0_X))
        c0 = np.ascontiguousarray(h0_X.T @ Yres)
        hb = np.ascontiguousarray(np.asarray(h0_betas, dtype=np.float64).reshape(q0))
        slab = result_empty((4, snp_count))
        out = {k: slab[i] for i, k in enumerate(('ps', 'f_stats', 'rss', 'var_perc'))}
        out['betas'] = result_empty((snp_count, q0 + 1))
        self._ck(self.lib.mmg_emmax_scan_betas_f64(self.h, R.handle, _ptr(V), q0, _ptr(Ainv), _ptr(c0), float(Yres @ Yres), _ptr(hb),
                                                   float(h0_rss), float(n_p), impl_id(impl), snp_begin, snp_count, _ptr(out['ps']),
                                                   _ptr(out['f_stats']), _ptr(out['rss']), _ptr(out['var_perc']), _ptr(out['betas'])))
        return out

    def emmax_scan_rows(self, xs, R, V, h0_rss, n_p, want_dots=False, want_stats=True, **_ignored):
        """emmax_scan for REAL-VALUED genotype rows xs [m x n] (imputed dosages): FP64 tensor-core path
        (mmg_emmax_scan_rows_f64); the rows pass through the device in chunks, the resident block is untouched."""
        xs = np.ascontiguousarray(xs, dtype=np.float64)
        m, n = xs.shape
        V = np.ascontiguousarray(np.asarray(V, dtype=np.float64))
        if V.ndim == 1:
            V = V.reshape(1, -1)
        nv = V.shape[0]
        assert V.shape[1] == R.shape[0] and R.shape[1] == n, (V.shape, R.shape, xs.shape)
        keys = (('ps', 'f_stats', 'rss', 'var_perc') if want_stats else ()) + ('xx',)
        slab = result_empty((len(keys), m))
        out = {k: slab[i] for i, k in enumerate(keys)}
        if want_dots:
            out['dots'] = result_empty((m, nv))
        self._ck(self.lib.mmg_emmax_scan_rows_f64(self.h, R.handle, _ptr(V), nv, float(h0_rss), float(n_p), _ptr(xs), m, n,
                                                  _ptr(out.get('ps')), _ptr(out.get('f_stats')), _ptr(out.get('rss')),
                                                  _ptr(out.get('var_perc')), _ptr(out['xx']), _ptr(out.get('dots'))))
        return out

    def emmax_scan_quad(self, A, v, h0_rss, n_p, snp_begin=0, snp_count=None):
        """The int8 scan given the quadratic form A = R'R (DeviceMatrix, lower triangle) and v = R'y~ (length n)."""
        m, n = self.snps_shape()
        if snp_count is None:
            snp_count = m - snp_begin
        v = np.ascontiguousarray(np.asarray(v, dtype=np.float64).reshape(-1))
        assert v.shape[0] == n and A.shape == (n, n), (v.shape, A.shape, n)
        keys = ('ps', 'f_stats', 'rss', 'var_perc', 'xx')
        slab = result_empty((len(keys), snp_count))
        out = {k: slab[i] for i, k in enumerate(keys)}
        self._ck(self.lib.mmg_emmax_scan_quad_f64(self.h, A.handle, _ptr(v), float(h0_rss), float(n_p), snp_begin, snp_count,
                                                  _ptr(out['ps']), _ptr(out['f_stats']), _ptr(out['rss']), _ptr(out['var_perc']),
                                                  _ptr(out['xx'])))
        return out

    def emmax_scan_quad_dev(self, A, v, h0_rss, n_p, packed=True, a_err=0.0, snp_begin=0, snp_count=None, out=None):
        """The int8 scan with everything on the device: A (packed blocks or dense), v = R'y~ (DeviceMatrix of n values); the
        results stay in `out`, a [5 x >= snp_count] DeviceMatrix with rows ps, f_stats, rss, var_perc, xx."""
        m, n = self.snps_shape()
        if snp_count is None:
            snp_count = m - snp_begin
        if out is None:
            out = DeviceMatrix(self, 5, snp_count)
        self._ck(self.lib.mmg_emmax_scan_quad_dev(self.h, A.handle, int(bool(packed)), float(a_err), v.handle if v is not None else 0, float(h0_rss), float(n_p),
                                                  int(snp_begin), int(snp_count), out.handle))
        return out

    def emmax_scan_multi(self, Rs, V, h0_rss, n_p, snp_begin=0, snp_count=None, want=('ps', 'f_stats', 'rss', 'var_perc')):
        """T phenotypes in one launch: Rs = list of T rotations (DeviceMatrix), V [T x n_out], h0_rss [T].
        Returns a dict of [T x snp_count] arrays."""
        m, n = self.snps_shape()
        if snp_count is None:
            snp_count = m - snp_begin
        T = len(Rs)
        V = np.ascontiguousarray(np.asarray(V, dtype=np.float64)).reshape(T, -1)
        assert V.shape[1] == Rs[0].shape[0], (V.shape, Rs[0].shape)
        h0 = np.ascontiguousarray(np.asarray(h0_rss, dtype=np.float64).reshape(T))
        handles = np.array([r.handle for r in Rs], dtype=np.int64)
        out = {k: result_empty((T, snp_count)) for k in want}
        self._ck(self.lib.mmg_emmax_scan_multi_f64(self.h, _ptr(handles), T, _ptr(V), _ptr(h0), float(n_p), snp_begin, snp_count,
                                                   _ptr(out.get('ps')), _ptr(out.get('f_stats')), _ptr(out.get('rss')),
                                                   _ptr(out.get('var_perc')), _ptr(out.get('xx'))))
        return out

    def emmax_scan_shared(self, U, Ext, W, q0, h0_rss, n_p, snp_begin=0, snp_count=None, want=('ps', 'f_stats', 'rss', 'var_perc')):
        """T phenotypes on one eigenbasis with ONE rotation per SNP (mmg_emmax_scan_shared_f64): U = eig_L vectors (rows),
        Ext [T (1 + q0) x n] = per phenotype v_t and the q0 rows c_tj, W [T x n] = 1 / (lambda + delta_t).
        Returns a dict of [T x snp_count] arrays plus 'info' (planes, certified bounds, kernel ms)."""
        m, n = self.snps_shape()
        if snp_count is None:
            snp_count = m - snp_begin
        W = np.ascontiguousarray(W, dtype=np.float64)
        T = W.shape[0]
        assert W.shape == (T, n) and Ext.shape == (T * (1 + q0), n) and U.shape == (n, n), (W.shape, Ext.shape, U.shape)
        h0 = np.ascontiguousarray(np.asarray(h0_rss, dtype=np.float64).reshape(T))
        out = {k: result_empty((T, snp_count)) for k in want}
        info = np.zeros(5)
        self._ck(self.lib.mmg_emmax_scan_shared_f64(self.h, U.handle, Ext.handle, _ptr(W), T, int(q0), _ptr(h0), float(n_p), snp_begin, snp_count,
                                                    _ptr(out.get('ps')), _ptr(out.get('f_stats')), _ptr(out.get('rss')),
                                                    _ptr(out.get('var_perc')), _ptr(out.get('xx')), _ptr(info)))
        out['info'] = {'planes': int(info[0]), 'rho_xx': info[1], 'rho_xy': info[2], 'rotation_ms': info[3], 'contraction_ms': info[4]}
        self.last_shared_info = out['info']
        return out

    def emmax_perm_scan(self, R, Wt, ratio, centre=True, impl=IMPL_AUTO, snp_begin=0, snp_count=None):
        m, n = self.snps_shape()
        if snp_count is None:
            snp_count = m - snp_begin
        assert ratio.dtype == np.float64 and ratio.shape[0] == Wt.shape[0]
        self._ck(self.lib.mmg_emmax_perm_scan_f64(self.h, R.handle, Wt.handle, int(bool(centre)), impl_id(impl), snp_begin,
                                                  snp_count, _ptr(ratio)))
        return ratio

    def f_sf(self, f, dfn, dfd):
        f = np.ascontiguousarray(f, dtype=np.float64)
        out = np.empty_like(f)
        self._ck(self.lib.mmg_f_sf_f64(self.h, _ptr(f), f.size, float(dfn), float(dfd), _ptr(out)))
        return out


def _as_int8(a):
    """Genotypes must be small integers for the tensor-core paths (the reference casts to int8 itself,
    kinship.py:31)."""
    a = np.asarray(a)
    if a.dtype == np.int8:
        return a
    if a.dtype.kind in 'iub':
        if a.size and (a.min() < -128 or a.max() > 127):
            raise ValueError('genotype values outside the int8 range')
        return a.astype(np.int8)
    if a.dtype.kind == 'f':
        r = np.rint(a)
        if not np.array_equal(r, a):
            raise TypeError('mixmogam_b200 scans integer genotype codes; got non-integral values '
                            '(there is no CPU fallback for real-valued dosages)')
        return _as_int8(r.astype(np.int64))
    raise TypeError('unsupported genotype dtype %r' % a.dtype)


class ResidentSnps(object):
    """Stands for "the genotype block that is resident on the device right now" wherever the API takes `snps`: readers that
    stream a file straight into HBM (hdf5_data.stream_snps) hand this to the scan instead of a host array."""

    def __init__(self, m, n):
        self.shape = (int(m), int(n))

    def __len__(self):
        return self.shape[0]


class PackedGenotypes(object):
    """Genotype codes 0..3 held at 2 bits each, SNP-major: `packed` is uint8 [m x >= ceil(n/4)], code j of a row in bits
    2 (j % 4), 2 (j % 4) + 1 of byte j // 4 (the layout of mmg_host_pack2).  Accepted wherever the API takes `snps`
    (calc_ibs_kinship, emmax, LinearMixedModel.emmax_f_test, ...): n / 4 bytes per SNP cross PCIe (SURVEY 8d's minimum) and the
    device expands them into its resident int8 block.  `len()` is the number of SNPs, `[i]` / `unpack()` give int8 rows back.
    A read-only `packed` array (the default of pack_genotypes) lets the device copy be reused across calls."""

    def __init__(self, packed, n):
        packed = np.asarray(packed)
        if packed.dtype != np.uint8 or packed.ndim != 2 or not packed.flags.c_contiguous or packed.shape[1] < (n + 3) // 4:
            raise ValueError('packed genotypes must be a C-contiguous uint8 [m x >= ceil(n/4)] array')
        self.packed = packed
        self.shape = (packed.shape[0], int(n))

    def __len__(self):
        return self.shape[0]

    def _key(self):
        return ('packed2', self.packed.ctypes.data, self.shape) if not self.packed.flags.writeable else None

    def unpack(self, rows=None):
        p = self.packed if rows is None else self.packed[rows]
        p = p.reshape(-1, self.packed.shape[1])
        out = np.empty((p.shape[0], 4 * p.shape[1]), dtype=np.int8)
        for k in range(4):
            out[:, k::4] = (p >> (2 * k)) & 3
        return out[:, :self.shape[1]]

    def __getitem__(self, i):
        if isinstance(i, (int, np.integer)):
            return self.unpack(slice(i, i + 1))[0]
        if isinstance(i, slice):
            return PackedGenotypes(self.packed[i], self.shape[1])
        return self.unpack(i)


def pack_genotypes(snps, threads=None, freeze=True):
    """int8 genotype codes 0..3 ([m x n] array or list of rows) -> PackedGenotypes in page-locked memory (mmg_host_pack2, all host
    threads).  freeze: mark the packed array read-only so that its device copy is kept between calls."""
    a = np.asarray(snps)
    if a.dtype != np.int8:
        a = _as_int8(a)
    a = np.ascontiguousarray(a)
    m, n = a.shape
    ld = ((n + 3) // 4 + 15) // 16 * 16
    try:
        packed = pinned_empty((m, ld), np.uint8)
    except MmgError:                      # no CUDA device in this process (packing is host-only work): ordinary memory
        packed = np.empty((m, ld), dtype=np.uint8)
    lib = load_library()
    rc = lib.mmg_host_pack2(_ptr(a), m, n, n, _ptr(packed), ld, int(threads or lib.mmg_host_threads_default()))
    if rc != 0:
        raise ValueError('genotype codes outside 0..3 cannot be packed to 2 bits')
    if freeze:
        packed.flags.writeable = False
    return PackedGenotypes(packed, n)


def real_valued(snps):
    """The genotypes as one FP64 [m x n] array if they hold non-integral values (imputed dosages), else None (integer codes
    go to the int8 tensor-core paths)."""
    if isinstance(snps, (PackedGenotypes, ResidentSnps)):
        return None
    if isinstance(snps, np.ndarray):
        if snps.dtype.kind != 'f':
            return None
        a = snps
    else:
        if len(snps) == 0 or all(getattr(r, 'dtype', None) is not None and r.dtype.kind in 'iub' for r in snps):
            return None
        a = np.asarray(snps)
        if a.dtype.kind != 'f':
            return None
    if a.ndim != 2:
        raise ValueError('genotypes must be an (m, n) array or a list of m rows')
    if np.array_equal(np.rint(a), a) and (a.size == 0 or (a.min() >= -128 and a.max() <= 127)):
        return None
    if not np.all(np.isfinite(a)):
        raise ValueError('genotypes contain NaN / inf')
    return np.ascontiguousarray(a, dtype=np.float64)


def resident(snps):
    """Marks a genotype array (or every row of a list of rows) read-only and returns it: the opt-in that lets the library
    keep its device copy across calls (Context.ensure_snps).  Make a copy, or set flags.writeable back, to edit it."""
    if isinstance(snps, np.ndarray):
        snps.flags.writeable = False
    else:
        for r in snps:
            r.flags.writeable = False
    return snps


_default_ctx = {}
_default_lock = threading.Lock()


def get_context(device=None):
    """Process-wide context per device (LOCAL_RANK / MMG_DEVICE select the default device)."""
    if device is None:
        device = int(os.environ.get('MMG_DEVICE', os.environ.get('LOCAL_RANK', '0')))
    with _default_lock:
        ctx = _default_ctx.get(device)
        if ctx is None or ctx.h is None:
            ctx = Context(device)
            _default_ctx[device] = ctx
    return ctx


class _PinnedBlock(object):
    """Owner of one page-locked host allocation; exposes it to numpy through __array_interface__ so that every
    array (and view) built on it keeps it alive.  When the last view dies the block goes back to a small pool:
    cudaHostAlloc of a 0.8 GB kinship costs more than copying it."""
    _pool = {}                      # nbytes -> [ptr]
    _pool_lock = threading.Lock()
    _POOL_MAX_PER_SIZE = 2          # large blocks (a 0.8 GB kinship); per-SNP result vectors keep up to 8 (see __del__)

    def __init__(self, nbytes):
        self.nbytes = int(nbytes)
        with _PinnedBlock._pool_lock:
            lst = _PinnedBlock._pool.get(self.nbytes)
            self.ptr = lst.pop() if lst else None
        if self.ptr is None:
            lib = load_library()
            p = C.c_void_p(0)
            rc = lib.mmg_host_alloc(C.byref(p), self.nbytes)
            if rc != 0:
                raise MmgError(rc, lib.mmg_last_error(None).decode())
            self.ptr = p.value
        self.__array_interface__ = {'shape': (self.nbytes,), 'typestr': '|u1', 'data': (self.ptr, False), 'version': 3}

    def __del__(self):
        try:
            ptr, self.ptr = self.ptr, None
            if ptr is None:
                return
            with _PinnedBlock._pool_lock:
                lst = _PinnedBlock._pool.setdefault(self.nbytes, [])
                if len(lst) < (8 if self.nbytes <= (64 << 20) else _PinnedBlock._POOL_MAX_PER_SIZE):
                    lst.append(ptr)
                    return
            load_library().mmg_host_free(C.c_void_p(ptr))
        except Exception:
            pass


def pinned_empty(shape, dtype=np.int8):
    """numpy array backed by page-locked host memory (mmg_host_alloc) for full-rate PCIe copies.  The memory is
    released (to a pool) when the array and all its views are garbage collected."""
    dtype = np.dtype(dtype)
    nbytes = max(1, int(np.prod(shape)) * dtype.itemsize)
    blk = _PinnedBlock(nbytes)
    return np.asarray(blk)[:int(np.prod(shape)) * dtype.itemsize].view(dtype).reshape(shape)


def result_empty(shape, dtype=np.float64):
    """Output buffer of a device call: page-locked (pooled) once it is large enough for the copy rate to matter --
    a pageable 8 MB per-SNP vector comes back at ~5 GB/s, a page-locked one at the PCIe rate."""
    dtype = np.dtype(dtype)
    nbytes = int(np.prod(shape)) * dtype.itemsize
    if nbytes < (1 << 20):
        return np.empty(shape, dtype=dtype)
    if nbytes <= (64 << 20):
        return pinned_empty(shape, dtype)
    # Large buffers: cudaHostAlloc runs at ~0.4 s / GB -- 0.55 of the 0.62 s of a one-off 199-phenotype x 214k-SNP batch, 3 s for
    # the 6.4 GB of a 1M-SNP one -- which only pays off when the block comes back from the pool.  The FIRST request of a size is
    # served from pageable memory (~0.1 s / GB to copy into); a second request of the same size (a loop) gets the page-locked block,
    # and every later one re-uses it.  Nothing above 1 GB is page-locked.
    if nbytes <= (1 << 30):
        with _PinnedBlock._pool_lock:
            pooled = bool(_PinnedBlock._pool.get(max(1, nbytes)))
            seen = _large_seen.get(nbytes, 0)
            _large_seen[nbytes] = seen + 1
        if pooled or seen >= 1:
            return pinned_empty(shape, dtype)
    return np.empty(shape, dtype=dtype)


_large_seen = {}


def pinned_free(arr):
    """Kept for API compatibility: pinned arrays are reference counted now."""
    return None
