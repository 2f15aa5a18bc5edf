#!/bin/bash
# configs[0] (n = 198 x 214k SNPs) through the public API: wall time and host profile.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo build failed; tail -20 gpurun_out/build.log; exit 1; }
timeout 300 python tools/latency_config0.py > gpurun_out/latency_config0.txt 2>&1; echo "rc=$?"; head -60 gpurun_out/latency_config0.txt | cut -c1-200
