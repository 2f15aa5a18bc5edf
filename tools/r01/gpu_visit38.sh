#!/bin/bash
# chained accumulation in the GEMM core (int8 R'R: 7 drains per output tile instead of 28): tests of every tc_gemm user, launch list
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo build failed; tail -20 gpurun_out/build.log; exit 1; }
timeout 900 python -m pytest tests/test_gpu_reml_scan.py tests/test_gpu_kinship.py tests/test_gpu_hdf5.py -q -m gpu -k "not streamed_from_host" -p no:cacheprovider --timeout 300 > gpurun_out/tests_tc.log 2>&1
echo "tests rc=$?"; tail -3 gpurun_out/tests_tc.log
export MMG_PROFILE_RANGE=1
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
   --log-file gpurun_out/launches_1m.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/launches_bench.log 2>&1
echo "launch list rc=$?"; python tools/ncu_summary.py launches gpurun_out/launches_1m.csv 2>/dev/null | head -9 | cut -c1-110
unset MMG_PROFILE_RANGE
timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e > gpurun_out/bench_chain.json 2> gpurun_out/bench_chain.err
python -c "
import json; d=json.load(open('gpurun_out/bench_chain.json')); print('value %.0f ms/step %.1f scan_kernel %.2f stages %s'%(d['value'], d['ms_per_step'], d['roofline']['launch_ms'], {k: round(1e3*v,1) for k,v in d['stage_seconds_per_step'].items() if v}))"
