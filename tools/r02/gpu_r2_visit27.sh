#!/bin/bash
# Host profile of configs[2] (n = 198, 214k SNPs, T = 199 phenotypes through emmax_multi).
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo build failed; tail -20 gpurun_out/build.log; exit 1; }
timeout 900 python -m cProfile -s cumulative tools/bench_configs.py --config 2 --single 2 > gpurun_out/config2_profile.txt 2>&1; echo "rc=$?"; grep -n '"config"' gpurun_out/config2_profile.txt | cut -c1-900; grep -A40 "Ordered by" gpurun_out/config2_profile.txt | cut -c1-170
