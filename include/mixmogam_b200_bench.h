/*
 * mixmogam_b200 -- pipe-rate microbenchmarks (libmixmogam_b200_bench.so).
 *
 * Diagnostics only: the product library (libmixmogam_b200.so, mixmogam_b200.h) carries no benchmark kernels.  This library
 * borrows a context created by mmg_create (its stream, scratch buffer and events) and must be built from the same sources.
 * bench.py uses it for the denominators of its roofline (profiles/PEAKS_int8_fp64.json).
 */
#ifndef MIXMOGAM_B200_BENCH_H
#define MIXMOGAM_B200_BENCH_H

#include "mixmogam_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* raw pipe rates measured with CUDA events: which =
 *   "dmma" / "dfma"            FP64 tensor (mma.sync m8n8k4) / FP64 FMA issue rate, TFLOP/s
 *   "imma_tcgen05", "imma_pair" [+ "_ldtm<k>"]   tcgen05 kind::i8 issue rate, smem-resident operands, TOP/s (burst: one launch)
 *   "imma_pair_sustained<ms>"  the same, repeated back to back for <ms> milliseconds (power-capped steady state), TOP/s
 *   "copy"                     device-to-device copy, GB/s
 *   "ldtm_w<4|8|16>_x<16|32>[_mma]"  SM cycles per 128 x 256 accumulator tile read back with tcgen05.ld
 *   "prepass_r<r>_u<u>_b<b>"   ms of the scan's linear pre-pass over the resident genotypes */
int mmg_microbench(mmg_ctx* ctx, const char* which, double* value);

#ifdef __cplusplus
}
#endif
#endif /* MIXMOGAM_B200_BENCH_H */
