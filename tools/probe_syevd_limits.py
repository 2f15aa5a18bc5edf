#!/usr/bin/env python
"""Which cuSOLVER symmetric eigensolver entry accepts n x n FP64 with n^2 >= 2^31?  Buffer-size queries only (no solve)."""
import ctypes as C
import sys
cs = C.CDLL('/usr/local/cuda/lib64/libcusolver.so')
rt = C.CDLL('/usr/local/cuda/lib64/libcudart.so')
h = C.c_void_p(); print('create', cs.cusolverDnCreate(C.byref(h)))
prm = C.c_void_p(); print('params', cs.cusolverDnCreateParams(C.byref(prm)))
CUDA_R_64F, VEC, LOWER = 1, 1, 0
dA = C.c_void_p(); dW = C.c_void_p()
rt.cudaMalloc(C.byref(dA), C.c_size_t(1 << 20)); rt.cudaMalloc(C.byref(dW), C.c_size_t(1 << 20))
for n in [int(a) for a in sys.argv[1:]] or [10000, 32768, 40000, 46340, 46341, 47000, 50000, 65536]:
    wd, wh = C.c_size_t(0), C.c_size_t(0)
    st = cs.cusolverDnXsyevd_bufferSize(h, prm, VEC, LOWER, C.c_int64(n), CUDA_R_64F, dA, C.c_int64(n), CUDA_R_64F, dW, CUDA_R_64F, C.byref(wd), C.byref(wh))
    out = ['n=%d Xsyevd st=%d dev=%.2f GB host=%d' % (n, st, wd.value / 1e9, wh.value)]
    wd, wh = C.c_size_t(0), C.c_size_t(0); meig = C.c_int64(0); vl = C.c_double(0); vu = C.c_double(0)
    st = cs.cusolverDnXsyevdx_bufferSize(h, prm, VEC, 0, LOWER, C.c_int64(n), CUDA_R_64F, dA, C.c_int64(n), C.byref(vl), C.byref(vu), C.c_int64(0), C.c_int64(0),
                                         C.byref(meig), CUDA_R_64F, dW, CUDA_R_64F, C.byref(wd), C.byref(wh))
    out.append('Xsyevdx st=%d dev=%.2f GB' % (st, wd.value / 1e9))
    wd, wh = C.c_size_t(0), C.c_size_t(0)
    st = cs.cusolverDnXsyevBatched_bufferSize(h, prm, VEC, LOWER, C.c_int64(n), CUDA_R_64F, dA, C.c_int64(n), CUDA_R_64F, dW, CUDA_R_64F, C.byref(wd), C.byref(wh), C.c_int64(1))
    out.append('XsyevBatched st=%d dev=%.2f GB' % (st, wd.value / 1e9))
    lw = C.c_int(0)
    st = cs.cusolverDnDsyevd_bufferSize(h, VEC, LOWER, C.c_int(n), dA, C.c_int(n), dW, C.byref(lw))
    out.append('Dsyevd st=%d lwork=%d' % (st, lw.value))
    info = C.c_void_p(); cs.cusolverDnCreateSyevjInfo(C.byref(info))
    lw = C.c_int(0)
    st = cs.cusolverDnDsyevj_bufferSize(h, VEC, LOWER, C.c_int(n), dA, C.c_int(n), dW, C.byref(lw), info)
    out.append('Dsyevj st=%d lwork=%d' % (st, lw.value))
    print(' | '.join(out), flush=True)
