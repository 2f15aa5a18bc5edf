#!/bin/bash
# validation: whole gpu suite, smoke, both bench arms
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo build failed; tail -20 gpurun_out/build.log; exit 1; }
timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider --timeout 600 > gpurun_out/tests_gpu.log 2>&1
echo "tests rc=$?"; tail -15 gpurun_out/tests_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; cut -c1-300 gpurun_out/bench_ref.json
timeout 900 python bench.py > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; echo "ours rc=$?"; cat gpurun_out/bench_ours.json; tail -5 gpurun_out/bench_ours.err
