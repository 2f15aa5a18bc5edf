#!/bin/bash
# round-end evidence, part A: whole gpu suite, smoke, ncu launch list + full captures (scan, int8 R'R, Gram)
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo build failed; tail -20 gpurun_out/build.log; exit 1; }
timeout 1200 python -m pytest tests -q -m gpu -p no:cacheprovider --timeout 600 > gpurun_out/tests_gpu.log 2>&1
echo "tests rc=$?"; tail -4 gpurun_out/tests_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
export MMG_PROFILE_RANGE=1
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
   --log-file gpurun_out/launches_1m.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/launches_bench.log 2>&1
echo "launch list rc=$?"
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:scan_quad_kernel -c 2 \
   -o gpurun_out/prof_scan_1m -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/prof_scan_1m.log 2>&1
echo "full capture scan rc=$?"
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:OzakiEpi -c 1 \
   -o gpurun_out/prof_ozaki_1m -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/prof_ozaki_1m.log 2>&1
echo "full capture R'R rc=$?"
