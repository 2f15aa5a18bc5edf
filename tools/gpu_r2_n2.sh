#!/bin/bash
# 2 GPUs: the sharded-path test (tests/test_gpu_multi.py) and the 2-GPU bench line at HEAD.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo build failed; tail -20 gpurun_out/build.log; exit 1; }
timeout 600 python -m pytest tests/test_gpu_multi.py -q -m gpu -p no:cacheprovider > gpurun_out/t_multi.log 2>&1; echo "t_multi rc=$?"; tail -4 gpurun_out/t_multi.log
NS=2 STEPS=8 bash tools/r02/gpu_r2_n8.sh
