#!/bin/bash
# round-end evidence, part B: int8 R'R block order check, both bench arms, launch list
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo build failed; tail -20 gpurun_out/build.log; exit 1; }
timeout 300 python -m pytest tests/test_gpu_reml_scan.py -x -q -m gpu -k "quad_form or golden or certified" -p no:cacheprovider --timeout 200 > gpurun_out/tests_quad.log 2>&1
echo "quad tests rc=$?"; tail -2 gpurun_out/tests_quad.log
timeout 300 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; cut -c1-250 gpurun_out/bench_ref.json
timeout 600 python bench.py > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; echo "ours rc=$?"; tail -2 gpurun_out/bench_ours.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_ours.json'))
print('value %.0f ms %.1f | e2e %.0f ms %.1f lanes %s'%(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e'].get('h2d_lanes')))
print('resident', {k: round(1e3*v,1) for k,v in d['stage_seconds_per_step'].items() if v})
print('e2e     ', {k: round(1e3*v,1) for k,v in d['e2e']['stage_seconds_per_step'].items() if v})
print('roofline', d['roofline'])
print('cpu', d['cpu_baseline']['value'], d['clocks'])
PY
export MMG_PROFILE_RANGE=1
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
   --log-file gpurun_out/launches_1m.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/launches_bench.log 2>&1
echo "launch list rc=$?"; python tools/ncu_summary.py launches gpurun_out/launches_1m.csv 2>/dev/null | head -9 | cut -c1-110
