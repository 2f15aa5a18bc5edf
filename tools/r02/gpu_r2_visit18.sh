#!/bin/bash
# configs[3] at the full n = 50 000 x 500 000 SNPs on one B200 with stand-in eigenbases (cuSOLVER syevd stops at n = 32768):
# the memory plan of kinship + REML + scan at that size.  A small run of the same code path first.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo build failed; tail -20 gpurun_out/build.log; exit 1; }
timeout 300 python tools/bench_configs.py --config 3 --indivs 3000 --snps 20000 --standin-eigen > gpurun_out/r02_config3_small.json 2> gpurun_out/r02_config3_small.err; echo "small rc=$?"; tail -c 700 gpurun_out/r02_config3_small.json; tail -3 gpurun_out/r02_config3_small.err
timeout 1200 python tools/bench_configs.py --config 3 --standin-eigen > gpurun_out/r02_config3_n50k.json 2> gpurun_out/r02_config3_n50k.err; echo "n50k rc=$?"; tail -c 1500 gpurun_out/r02_config3_n50k.json; tail -5 gpurun_out/r02_config3_n50k.err
nvidia-smi --query-gpu=memory.used,memory.total --format=csv,noheader
