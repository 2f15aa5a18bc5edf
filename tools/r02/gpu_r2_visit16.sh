#!/bin/bash
# Pack kernel with bit-operation thermometer planes: kinship tests, pack / Gram times in the bench step.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo build failed; tail -20 gpurun_out/build.log; exit 1; }
timeout 600 python -m pytest tests/test_gpu_kinship.py -q -m gpu -p no:cacheprovider > gpurun_out/t_kin.log 2>&1; echo "t_kin rc=$?"; tail -4 gpurun_out/t_kin.log
timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_pack.json 2> gpurun_out/bench_pack.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_pack.json').read().strip().splitlines()[-1])
print('value %.0f ms/step %.1f gram %.2f' % (d['value'], d['ms_per_step'], d['kinship']['gram_ms']), {k: round(1e3*v,2) for k,v in d['stage_seconds_per_step'].items() if v}, d['clocks']['sm_mhz'])
PY
