#!/bin/bash
# base-256 digit planes + streamed host Gram: new tests first, whole gpu suite, bench, then ncu evidence at the bench size
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo build failed; tail -20 gpurun_out/build.log; exit 1; }
timeout 400 python -m pytest tests/test_gpu_kinship.py tests/test_gpu_reml_scan.py -x -q -m gpu -k "streamed or certified or golden" -p no:cacheprovider --timeout 200 > gpurun_out/tests_new.log 2>&1
echo "new tests rc=$?"; tail -4 gpurun_out/tests_new.log
timeout 1200 python -m pytest tests -q -m gpu -p no:cacheprovider --timeout 600 --durations=12 > gpurun_out/tests_gpu.log 2>&1
echo "tests rc=$?"; tail -22 gpurun_out/tests_gpu.log
timeout 600 python bench.py > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; echo "ours rc=$?"; cat gpurun_out/bench_ours.json; tail -3 gpurun_out/bench_ours.err
export MMG_PROFILE_RANGE=1
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
   --log-file gpurun_out/launches_1m.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/launches_bench.log 2>&1
echo "launch list rc=$?"
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:scan_quad_kernel -c 2 \
   -o gpurun_out/prof_scan_1m -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/prof_scan_1m.log 2>&1
echo "full capture scan rc=$?"
ls -la gpurun_out | tail -8
