#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo build failed; tail -20 gpurun_out/build.log; exit 1; }
timeout 900 python -m pytest tests/test_gpu_kinship.py -q -m gpu -p no:cacheprovider > gpurun_out/t_kin.log 2>&1; echo "t_kin rc=$?"; tail -4 gpurun_out/t_kin.log
