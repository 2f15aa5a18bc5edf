#!/bin/bash
# Staged re-pitch upload of short genotype rows: tests that upload n = 198 / 400 blocks, configs[0] wall time.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo build failed; tail -20 gpurun_out/build.log; exit 1; }
timeout 900 python -m pytest tests/test_gpu_reml_scan.py tests/test_gpu_reference_pin.py tests/test_gpu_hdf5.py tests/test_gpu_emma.py -q -m gpu -p no:cacheprovider > gpurun_out/t_scan.log 2>&1; echo "t_scan rc=$?"; tail -4 gpurun_out/t_scan.log
timeout 300 python tools/latency_config0.py > gpurun_out/latency_config0.txt 2>&1; echo "rc=$?"; head -12 gpurun_out/latency_config0.txt | cut -c1-330
