#!/bin/bash
# Host profile of configs[4] (hdf5_data.run_emmax_perm, n = 5000 x 1M SNPs x 1000 permutations) + the Jacobi test.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo build failed; tail -20 gpurun_out/build.log; exit 1; }
timeout 300 python -m pytest tests/test_gpu_reml_scan.py -q -m gpu -p no:cacheprovider -k "jacobi or syevd" > gpurun_out/t_jac.log 2>&1; echo "t_jac rc=$?"; tail -3 gpurun_out/t_jac.log
timeout 900 python -m cProfile -s cumulative tools/bench_configs.py --config 4 > gpurun_out/config4_profile.txt 2>&1; echo "rc=$?"; grep -n "config" gpurun_out/config4_profile.txt | head -3; grep -A45 "Ordered by" gpurun_out/config4_profile.txt | cut -c1-170
