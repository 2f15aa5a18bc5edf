#!/bin/bash
# Gram tile order: blocked walk of the triangle (L2 reuse) against the column-by-column order; bit-exactness of the blocked order.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo build failed; tail -20 gpurun_out/build.log; exit 1; }
MMG_GRAM_BLOCKED=1 timeout 300 python -m pytest tests/test_gpu_kinship.py -q -m gpu -p no:cacheprovider -k "split_k or (gram_bit_exact and tcgen05 and 2048)" > gpurun_out/t_kin.log 2>&1; echo "t_kin rc=$?"; tail -3 gpurun_out/t_kin.log
for b in 1 0 1 0; do
  MMG_GRAM_BLOCKED=$b timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_blk$b.json 2> gpurun_out/bench_blk$b.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/bench_blk$b.json').read().strip().splitlines()[-1])
print('blocked=$b value %.0f ms/step %.1f gram %.2f frac %.3f clocks %s' % (d['value'], d['ms_per_step'], d['kinship']['gram_ms'], d['kinship']['frac'], d['clocks']['sm_mhz']))
PY
done
