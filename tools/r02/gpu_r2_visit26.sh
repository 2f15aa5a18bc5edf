#!/bin/bash
# Resident-block allocation re-used across shapes, cached stream ring, 1-D permutation shuffle: whole GPU suite, configs[4] timing.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo build failed; tail -20 gpurun_out/build.log; exit 1; }
timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider --deselect tests/test_gpu_full_size.py > gpurun_out/t_all.log 2>&1; echo "t_all rc=$?"; tail -5 gpurun_out/t_all.log
timeout 900 python tools/bench_configs.py --config 4 > gpurun_out/r02_config4.json 2> gpurun_out/r02_config4.err; echo "config4 rc=$?"; cat gpurun_out/r02_config4.json | cut -c1-900
