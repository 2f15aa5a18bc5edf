#!/bin/bash
# bench at a small and at the full size, then the ncu launch list of a short run
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 600 python -m pytest tests/test_gpu_reml_scan.py -q -m gpu -k "perm" -p no:cacheprovider --timeout 300 > gpurun_out/perm.log 2>&1; echo "perm rc=$?"; tail -3 gpurun_out/perm.log
timeout 600 python bench.py --snps 131072 --steps 2 --warmup 1 > gpurun_out/bench_small.json 2> gpurun_out/bench_small.err; echo "bench_small rc=$?"; cat gpurun_out/bench_small.json; tail -5 gpurun_out/bench_small.err
timeout 1500 python bench.py --steps 2 --warmup 1 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "bench_full rc=$?"; cat gpurun_out/bench_full.json; tail -5 gpurun_out/bench_full.err
