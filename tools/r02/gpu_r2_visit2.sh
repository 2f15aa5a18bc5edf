#!/bin/bash
# Round 2, visit 2: batched-EMMA / ML tests, the scan tests after the workspace change, host profile of a step.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo build failed; tail -20 gpurun_out/build.log; exit 1; }
run() { name=$1; shift; timeout 900 python -m pytest "$@" -q -m gpu -p no:cacheprovider -x > gpurun_out/$name.log 2>&1; echo "$name rc=$?"; tail -4 gpurun_out/$name.log; }
run t_emma tests/test_gpu_emma.py
run t_scan tests/test_gpu_reml_scan.py tests/test_gpu_reference_pin.py
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --profile-host gpurun_out/host_profile.txt > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_n1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n1.json').read().strip().splitlines()[-1])
print('N=1 value %.0f ms/step %.1f scan_kernel %.1f frac %.3f e2e %.0f (%.1f ms)'%(d['value'], d['ms_per_step'], d['roofline']['launch_ms'], d['roofline']['frac'], d['e2e']['value'], d['e2e']['ms_per_step']))
print('stages', {k: round(1e3*v,2) for k,v in d['stage_seconds_per_step'].items() if v})
PY
head -70 gpurun_out/host_profile.txt
