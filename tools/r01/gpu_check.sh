#!/bin/bash
# One GPU-box visit: full gpu test-suite, smoke, full-size bench, reference arm, ncu launch list + full capture of the scan kernel.
mkdir -p gpurun_out
nvidia-smi > gpurun_out/nvsmi.txt 2>&1
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
( time timeout 1500 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider --durations=8 ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest gpu rc=$?"; tail -15 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "bench_full rc=$?"; cat gpurun_out/bench_full.json; tail -5 gpurun_out/bench_full.err
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; cat gpurun_out/bench_ref.json
export MMG_PROFILE_RANGE=1
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
   --log-file gpurun_out/launches.csv python bench.py --snps 131072 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/launches_bench.log 2>&1
echo "launch list rc=$?"
timeout 1200 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:QuadEpi -c 1 \
   -o gpurun_out/prof_quad -f python bench.py --snps 131072 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/prof_quad.log 2>&1
echo "full capture quad rc=$?"
