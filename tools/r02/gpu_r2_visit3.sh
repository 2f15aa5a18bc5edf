#!/bin/bash
# Round 2, visit 3: the whole GPU suite except the full-size test, smoke, 1-GPU bench.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo build failed; tail -20 gpurun_out/build.log; exit 1; }
timeout 1200 python -m pytest tests -q -m gpu -p no:cacheprovider -x --deselect tests/test_gpu_full_size.py > gpurun_out/t_all.log 2>&1; echo "t_all rc=$?"; tail -6 gpurun_out/t_all.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_n1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n1.json').read().strip().splitlines()[-1])
print('N=1 value %.0f ms/step %.1f scan_kernel %.1f frac %.3f e2e %.0f (%.1f ms)'%(d['value'], d['ms_per_step'], d['roofline']['launch_ms'], d['roofline']['frac'], d['e2e']['value'], d['e2e']['ms_per_step']))
print('stages', {k: round(1e3*v,2) for k,v in d['stage_seconds_per_step'].items() if v})
print('e2e stages', {k: round(1e3*v,2) for k,v in d['e2e']['stage_seconds_per_step'].items() if v})
PY
