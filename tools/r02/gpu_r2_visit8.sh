#!/bin/bash
# Round 2, visit 8: ncu --set full of the shared-rotation kernels (rotation GEMM, contraction), then the batch timing again.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo build failed; tail -20 gpurun_out/build.log; exit 1; }
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"tc_gemm_i8_kernel|scan_dmma_kernel|shared_finish" -s 4 -c 3 -o gpurun_out/prof_shared \
  python tools/bench_multi.py --indivs 10000 --snps 16384 --phenotypes 199 --single 0 --unshared 0 > gpurun_out/prof_shared.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/prof_shared.log
ncu -i gpurun_out/prof_shared.ncu-rep --page raw --csv > gpurun_out/prof_shared_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_shared.ncu-rep --page details > gpurun_out/prof_shared_details.txt 2>/dev/null
ncu -i gpurun_out/prof_shared.ncu-rep --page source --csv > gpurun_out/prof_shared_source.csv 2>/dev/null
timeout 600 python tools/bench_multi.py --indivs 10000 --snps 131072 --phenotypes 199 --single 1 --unshared 0 > gpurun_out/r02_multi.json 2> gpurun_out/r02_multi.err; echo "multi rc=$?"; tail -c 1800 gpurun_out/r02_multi.json; tail -5 gpurun_out/r02_multi.err
