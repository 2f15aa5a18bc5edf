#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo build failed; tail -20 gpurun_out/build.log; exit 1; }
timeout 600 python -m pytest tests/test_gpu_emma.py tests/test_gpu_reml_scan.py tests/test_gpu_reference_pin.py -q -m gpu -p no:cacheprovider > gpurun_out/t_scan.log 2>&1; echo "t_scan rc=$?"; tail -3 gpurun_out/t_scan.log
