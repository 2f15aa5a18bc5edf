#!/bin/bash
# CTA-pair form of the e2m1 Gram (gram_pair.cuh): kinship tests under a timeout (new barrier protocol), A/B against the multicast form.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo build failed; tail -20 gpurun_out/build.log; exit 1; }
timeout 120 python -m pytest tests/test_gpu_kinship.py -q -m gpu -p no:cacheprovider -x -k "gram_bit_exact and tcgen05 and not i8" > gpurun_out/t_kin0.log 2>&1; echo "t_kin0 rc=$?"; tail -5 gpurun_out/t_kin0.log
timeout 600 python -m pytest tests/test_gpu_kinship.py tests/test_gpu_reference_pin.py -q -m gpu -p no:cacheprovider > gpurun_out/t_kin.log 2>&1; echo "t_kin rc=$?"; tail -8 gpurun_out/t_kin.log
run() {  # name, env...
  name=$1; shift
  env "$@" timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err; echo "bench $name rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_$name.json').read().strip().splitlines()[-1])
    print('$name', 'value %.0f ms/step %.1f gram %.2f frac %.3f' % (d['value'], d['ms_per_step'], d['kinship']['gram_ms'], d['kinship']['frac']), {k: round(1e3*v,2) for k,v in d['stage_seconds_per_step'].items() if v}, d['clocks']['sm_mhz'])
except Exception as e:
    print('$name failed', e)
PY
}
run pair1 MMG_GRAM_PAIR=1
run pair0 MMG_GRAM_PAIR=0
run pair1_b MMG_GRAM_PAIR=1
