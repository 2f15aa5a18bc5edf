#!/bin/bash
# two-lane host upload: kinship tests, full-size test, bench with e2e (lanes reported), host pack rate
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo build failed; tail -20 gpurun_out/build.log; exit 1; }
nproc; lscpu | grep -E "Model name|Socket|Thread|Core" | head -5
timeout 600 python -m pytest tests/test_gpu_kinship.py tests/test_gpu_full_size.py -x -q -m gpu -p no:cacheprovider --timeout 300 > gpurun_out/tests_kin.log 2>&1
echo "kinship tests rc=$?"; tail -4 gpurun_out/tests_kin.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; echo "ours rc=$?"; tail -3 gpurun_out/bench_ours.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_ours.json'))
print('value %.0f ms %.1f | e2e %.0f ms %.1f lanes %s'%(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e'].get('h2d_lanes')))
print('resident', {k: round(1e3*v,1) for k,v in d['stage_seconds_per_step'].items() if v})
print('e2e     ', {k: round(1e3*v,1) for k,v in d['e2e']['stage_seconds_per_step'].items() if v})
PY
MMG_H2D_PACK=0 timeout 600 python bench.py --no-cpu-baseline --steps 2 --warmup 2 > gpurun_out/bench_nopack.json 2> gpurun_out/bench_nopack.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_nopack.json'))
print('MMG_H2D_PACK=0: e2e %.0f ms %.1f lanes %s'%(d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e'].get('h2d_lanes')))
PY
