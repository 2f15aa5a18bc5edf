#!/bin/bash
# Round 2, visit 5: the tests touched in this session, cuSOLVER size limit between 32768 and 40000, configs[3] at n = 32768.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo build failed; tail -20 gpurun_out/build.log; exit 1; }
timeout 900 python -m pytest tests/test_gpu_emma.py tests/test_gpu_hdf5.py -q -m gpu -p no:cacheprovider > gpurun_out/t_new.log 2>&1; echo "t_new rc=$?"; tail -8 gpurun_out/t_new.log
python tools/probe_syevd_limits.py 33000 34000 35000 36000 37000 38000 39000 > gpurun_out/syevd_limits.txt 2>&1; cat gpurun_out/syevd_limits.txt | cut -c1-60
timeout 1200 python tools/bench_configs.py --config 3 --indivs 32768 > gpurun_out/r02_config3_n32768.json 2> gpurun_out/r02_config3_n32768.err
echo "config 3 rc=$?"; tail -c 1500 gpurun_out/r02_config3_n32768.json; tail -3 gpurun_out/r02_config3_n32768.err
