#!/bin/bash
# Short FP64 scans split over the SMs: scan tests, single-SNP latency at n = 1500 / 10000.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo build failed; tail -20 gpurun_out/build.log; exit 1; }
timeout 900 python -m pytest tests/test_gpu_reml_scan.py tests/test_gpu_emma.py tests/test_gpu_reference_pin.py -q -m gpu -p no:cacheprovider > gpurun_out/t_scan.log 2>&1; echo "t_scan rc=$?"; tail -5 gpurun_out/t_scan.log
timeout 300 python tools/latency_single_snp.py 1500 4000 > gpurun_out/latency_n1500.txt 2>&1; echo "n1500 rc=$?"; grep "ms per call" gpurun_out/latency_n1500.txt
timeout 600 python tools/latency_single_snp.py 10000 20000 > gpurun_out/latency_n10000.txt 2>&1; echo "n10000 rc=$?"; grep "ms per call" gpurun_out/latency_n10000.txt
