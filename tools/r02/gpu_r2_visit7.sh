#!/bin/bash
# Round 2, visit 7: tests of the fused with_betas finish + cooperative wave barrier; ncu --set full of the shared-rotation kernels.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo build failed; tail -20 gpurun_out/build.log; exit 1; }
timeout 900 python -m pytest tests/test_gpu_reml_scan.py tests/test_gpu_emma.py tests/test_gpu_reference_pin.py -q -m gpu -p no:cacheprovider -x > gpurun_out/t_new.log 2>&1; echo "t_new rc=$?"; tail -6 gpurun_out/t_new.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"tc_gemm_i8_kernel|scan_dmma_kernel|shared_finish" -s 6 -c 3 -o gpurun_out/prof_shared \
  python tools/bench_multi.py --indivs 10000 --snps 16384 --phenotypes 199 --single 0 --unshared 0 > gpurun_out/prof_shared.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/prof_shared.log
ncu -i gpurun_out/prof_shared.ncu-rep --page raw --csv > gpurun_out/prof_shared_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_shared.ncu-rep --page details > gpurun_out/prof_shared_details.txt 2>/dev/null
ls -la gpurun_out/prof_shared*
