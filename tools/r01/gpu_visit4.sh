#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
for cs in 2 4 1; do
  MMG_SCAN_CLUSTER=$cs MMG_GRAM_CLUSTER=$cs timeout 600 python -m pytest tests/test_gpu_kinship.py tests/test_gpu_reml_scan.py -x -q -m gpu -k "tcgen05 or agree" -p no:cacheprovider --timeout 300 > gpurun_out/tests_cs$cs.log 2>&1
  echo "tests cs=$cs rc=$?"; tail -2 gpurun_out/tests_cs$cs.log
done
for cs in 1 2 4; do
  MMG_SCAN_CLUSTER=$cs MMG_GRAM_CLUSTER=$cs timeout 600 python bench.py --snps 262144 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/bench_cs$cs.json 2> gpurun_out/bench_cs$cs.err
  echo "bench cs=$cs rc=$?"; python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_cs$cs.json'))
    print('cs=$cs value %.0f scan_ms %.2f frac %.3f gram_ms %.2f clocks %s'%(d['value'], d['roofline']['launch_ms'], d['roofline']['frac'], d['kinship']['gram_ms'], d['clocks']))
except Exception as e: print('parse fail', e)
PY
done
