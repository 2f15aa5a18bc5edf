#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo build failed; tail -20 gpurun_out/build.log; exit 1; }
timeout 600 python -m pytest tests/test_gpu_kinship.py -q -m gpu -p no:cacheprovider > gpurun_out/t_kin.log 2>&1; echo "t_kin rc=$?"; tail -3 gpurun_out/t_kin.log
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_chk.json 2> gpurun_out/bench_chk.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_chk.json').read().strip().splitlines()[-1])
print('value %.0f ms/step %.1f' % (d['value'], d['ms_per_step']), d['kinship']['kernel'], d['kinship']['gram_ms'], d['roofline']['frac'], d['roofline'].get('frac_of_rate_with_digit_operands'))
PY
