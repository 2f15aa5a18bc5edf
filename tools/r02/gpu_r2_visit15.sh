#!/bin/bash
# Pre-packed rows copied contiguously: packed-input tests and the bench with its e2e variants.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo build failed; tail -20 gpurun_out/build.log; exit 1; }
timeout 600 python -m pytest tests/test_gpu_hdf5.py -q -m gpu -p no:cacheprovider > gpurun_out/t_hdf5.log 2>&1; echo "t_hdf5 rc=$?"; tail -4 gpurun_out/t_hdf5.log
timeout 900 python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/bench_packed.json 2> gpurun_out/bench_packed.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_packed.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_packed.json').read().strip().splitlines()[-1])
print('N=1 value %.0f ms/step %.1f e2e %.0f (%.1f ms)'%(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step']))
print('e2e stages', {k: round(1e3*v,2) for k,v in d['e2e']['stage_seconds_per_step'].items() if v})
for k,v in d['e2e'].get('other_host_buffers',{}).items(): print(' e2e', k, '%.0f /s %.1f ms' % (v['value'], v['ms_per_step']), {a: round(1e3*b,1) for a,b in v['stage_seconds_per_step'].items() if b})
PY
