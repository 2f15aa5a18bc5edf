#!/bin/bash
# measured raw-lane backlog in the two-lane upload + tuned pre-pass: streamed Gram tests, e2e bench
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo build failed; tail -20 gpurun_out/build.log; exit 1; }
timeout 600 python -m pytest tests/test_gpu_kinship.py -x -q -m gpu -k "streamed or split" -p no:cacheprovider --timeout 300 > gpurun_out/tests_kin.log 2>&1
echo "streamed tests rc=$?"; tail -3 gpurun_out/tests_kin.log
for th in 16 12; do
MMG_HOST_THREADS=$th timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_th$th.json 2> gpurun_out/bench_th$th.err; echo "threads=$th rc=$?"; tail -2 gpurun_out/bench_th$th.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_th$th.json'))
print('threads $th: value %.0f ms %.1f | e2e %.0f ms %.1f lanes %s'%(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e'].get('h2d_lanes')))
print('resident', {k: round(1e3*v,1) for k,v in d['stage_seconds_per_step'].items() if v})
print('e2e     ', {k: round(1e3*v,1) for k,v in d['e2e']['stage_seconds_per_step'].items() if v})
PY
done
