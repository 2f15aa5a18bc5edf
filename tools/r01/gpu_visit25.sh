#!/bin/bash
# TMEM read-back microbenchmark: cycles per 128x256 int32 tile for the epilogue pattern, 4/8/16 warps, x16/x32, with and without the MMA running
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo build failed; tail -20 gpurun_out/build.log; exit 1; }
timeout 300 python - <<'PY' 2>&1 | tee gpurun_out/microbench_ldtm.txt
from mixmogam_b200 import get_context
ctx = get_context(0)
print('# mmg_microbench ldtm_*: SM cycles per 128x256 int32 accumulator tile read back (tcgen05.ld 32x32b, double buffered, IMAD per element)')
for mma in ('', '_mma'):
    for w in (4, 8, 16):
        for x in (16, 32):
            name = 'ldtm_w%d_x%d%s' % (w, x, mma)
            print(name, '%.0f' % ctx.microbench(name), flush=True)
print('imma_tcgen05', '%.0f' % ctx.microbench('imma_tcgen05'))
PY
