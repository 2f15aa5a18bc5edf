#!/bin/bash
# pre-pass tuning variants on 262144 resident SNPs x 10000
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo build failed; tail -20 gpurun_out/build.log; exit 1; }
timeout 300 python - <<'PY' 2>&1 | tee gpurun_out/microbench_prepass.txt
import torch, bench
from mixmogam_b200 import get_context
ctx = get_context(0)
m, n = 262144, 10000
snps = bench.gen_genotypes_pinned(0, m, n, torch.device('cuda:0'))
ctx.ensure_snps(snps)
print('# snp_prepass_kernel<ROWS per warp, UNROLL, min blocks/SM>: ms for %d SNPs x %d (x %.2f for 1M)' % (m, n, 1e6 / m))
for name in ('r4_u4_b3', 'r4_u4_b1', 'r4_u2_b3', 'r4_u8_b1', 'r4_u4_b4', 'r4_u8_b4', 'r8_u4_b2', 'r8_u2_b3', 'r2_u4_b4', 'r2_u8_b4', 'r2_u16_b4', 'r1_u16_b4'):
    print(name, '%.3f ms' % ctx.microbench('prepass_' + name), flush=True)
print('dfma TFLOP/s', ctx.microbench('dfma'), 'copy GB/s', ctx.microbench('copy'))
PY
