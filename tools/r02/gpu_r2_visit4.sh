#!/bin/bash
# Round 2, visit 4: whole GPU suite except the full-size test; configs[3] at the largest n cuSOLVER's Xsyevd accepts (32768);
# n = 10 000 parity run against the line-faithful oracle.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo build failed; tail -20 gpurun_out/build.log; exit 1; }
timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider --deselect tests/test_gpu_full_size.py > gpurun_out/t_all.log 2>&1; echo "t_all rc=$?"; tail -15 gpurun_out/t_all.log
timeout 900 python tools/bench_configs.py --config 3 --indivs 32768 > gpurun_out/r02_config3_n32768.json 2> gpurun_out/r02_config3_n32768.err
echo "config 3 rc=$?"; tail -c 1500 gpurun_out/r02_config3_n32768.json; tail -3 gpurun_out/r02_config3_n32768.err
MMG_TEST_FULL_M=131072 MMG_TEST_ORACLE_FULL=1 timeout 1200 python -m pytest tests/test_gpu_full_size.py -q -s -p no:cacheprovider > gpurun_out/r02_parity_n10k.txt 2>&1
echo "parity rc=$?"; grep -a "oracle(double)\|passed\|failed" gpurun_out/r02_parity_n10k.txt
