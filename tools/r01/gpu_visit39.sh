#!/bin/bash
# round-end check at HEAD: whole gpu suite, smoke, bench
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo build failed; tail -20 gpurun_out/build.log; exit 1; }
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider --timeout 600 > gpurun_out/tests_gpu.log 2>&1
echo "tests rc=$?"; tail -3 gpurun_out/tests_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
timeout 400 python bench.py > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; echo "ours rc=$?"; tail -2 gpurun_out/bench_ours.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_ours.json'))
print('value %.0f ms %.1f | e2e %.0f ms %.1f lanes %s'%(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e'].get('h2d_lanes')))
print('roofline frac %.3f frac_issue %.3f scan %.1f ms S=%d rho %.3g launches %d'%(d['roofline']['frac'], d['roofline']['frac_of_issue_rate'], d['roofline']['launch_ms'], d['roofline']['slices'], d['roofline']['certified_rel_bound_xx'], d['gpu_launches']))
print('cpu', d['cpu_baseline']['value'], d['clocks'])
PY
