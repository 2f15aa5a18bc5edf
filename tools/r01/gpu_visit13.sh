#!/bin/bash
# ncu evidence for the new scan kernel at the full bench size: launch list + full captures (scan, gram)
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo build failed; tail -20 gpurun_out/build.log; exit 1; }
export MMG_PROFILE_RANGE=1
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
   --log-file gpurun_out/launches_1m.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/launches_bench.log 2>&1
echo "launch list rc=$?"
timeout 1200 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:scan_quad_kernel -c 2 \
   -o gpurun_out/prof_scan_1m -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/prof_scan_1m.log 2>&1
echo "full capture scan rc=$?"
timeout 1200 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:tc_gemm_i8_kernel -c 1 \
   -o gpurun_out/prof_gram_1m -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/prof_gram_1m.log 2>&1
echo "full capture gram rc=$?"
ls -la gpurun_out | tail -5
