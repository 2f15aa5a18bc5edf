#!/bin/bash
# Jacobi eigensolver for small matrices (n <= 512): whole GPU suite, configs[0] wall time with and without it.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo build failed; tail -20 gpurun_out/build.log; exit 1; }
timeout 300 python tools/latency_config0.py > gpurun_out/latency_config0.txt 2>&1; echo "jacobi rc=$?"; head -3 gpurun_out/latency_config0.txt | cut -c1-330
MMG_SYEVD_JACOBI_MAX=0 timeout 300 python tools/latency_config0.py > gpurun_out/latency_config0_syevd.txt 2>&1; echo "syevd rc=$?"; head -3 gpurun_out/latency_config0_syevd.txt | cut -c1-330
timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider --deselect tests/test_gpu_full_size.py > gpurun_out/t_all.log 2>&1; echo "t_all rc=$?"; tail -6 gpurun_out/t_all.log
