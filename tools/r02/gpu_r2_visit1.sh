#!/bin/bash
# Round 2, visit 1: build, the new / changed GPU tests, pipe rates, 1-GPU bench.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/nvsmi.txt 2>&1
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo build failed; tail -20 gpurun_out/build.log; exit 1; }
run() { name=$1; shift; timeout 900 python -m pytest "$@" -q -m gpu -p no:cacheprovider -x > gpurun_out/$name.log 2>&1; echo "$name rc=$?"; tail -4 gpurun_out/$name.log; }
run t_scan tests/test_gpu_reml_scan.py
run t_kin tests/test_gpu_kinship.py
run t_pin tests/test_gpu_reference_pin.py tests/test_gpu_hdf5.py
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 300 python tools/measure_peaks.py > gpurun_out/peaks.log 2>&1; echo "peaks rc=$?"; tail -30 gpurun_out/peaks.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_n1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n1.json').read().strip().splitlines()[-1])
print('N=1 value %.0f ms/step %.1f scan_kernel %.1f frac %.3f (burst %.3f) e2e %.0f (%.1f ms)'%(d['value'], d['ms_per_step'], d['roofline']['launch_ms'], d['roofline']['frac'], d['roofline']['frac_of_burst_issue_rate'] or 0, d['e2e']['value'], d['e2e']['ms_per_step']))
print('stages', {k: round(1e3*v,2) for k,v in d['stage_seconds_per_step'].items() if v})
print('e2e stages', {k: round(1e3*v,2) for k,v in d['e2e']['stage_seconds_per_step'].items() if v})
print('cpu', d['cpu_baseline'])
PY
timeout 900 python -m pytest tests/test_gpu_full_size.py -q -m gpu -p no:cacheprovider -x -s > gpurun_out/t_full.log 2>&1; echo "t_full rc=$?"; tail -5 gpurun_out/t_full.log
