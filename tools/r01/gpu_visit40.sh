#!/bin/bash
# LinearModel.fast_f_test on the GPU
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo build failed; tail -20 gpurun_out/build.log; exit 1; }
timeout 200 python -m pytest tests/test_gpu_reference_pin.py -x -q -m gpu -k "fast_f_test" -p no:cacheprovider --timeout 150 > gpurun_out/tests_fft.log 2>&1
echo "fast_f_test rc=$?"; tail -25 gpurun_out/tests_fft.log
