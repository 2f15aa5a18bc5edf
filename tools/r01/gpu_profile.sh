#!/bin/bash
# ncu evidence: launch list of the timed steps + full captures of the two tensor-core kernels
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
export MMG_PROFILE_RANGE=1
M=${M:-131072}
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
   --log-file gpurun_out/launches.csv python bench.py --snps $M --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/launches_bench.log 2>&1
echo "launch list rc=$?"
timeout 1200 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:tc_gemm_i8_kernel -c 2 \
   -o gpurun_out/prof_tc -f python bench.py --snps $M --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/prof_tc.log 2>&1
echo "full capture tc rc=$?"
timeout 1200 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:scan_dmma_kernel -c 1 \
   -o gpurun_out/prof_dmma -f python bench.py --snps 32768 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --scan-impl dmma > gpurun_out/prof_dmma.log 2>&1
echo "full capture dmma rc=$?"
ls -la gpurun_out
