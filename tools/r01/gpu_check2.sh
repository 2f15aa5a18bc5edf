#!/bin/bash
# new kernels: permutation scan (tcgen05), phenotype-batched scan, IBD int8 Gram; reference-pin tests; host profile; ncu of the scan
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
run() { name=$1; shift; timeout 900 python -m pytest "$@" -q -m gpu -p no:cacheprovider --timeout 600 > gpurun_out/$name.log 2>&1; echo "$name rc=$?"; tail -25 gpurun_out/$name.log; }
run t_perm tests/test_gpu_reml_scan.py -k "perm"
run t_multi tests/test_gpu_reml_scan.py -k "multi"
run t_ibd tests/test_gpu_kinship.py -k "ibd"
run t_pin tests/test_gpu_reference_pin.py
run t_rest tests/test_gpu_reml_scan.py tests/test_gpu_kinship.py tests/test_gpu_hdf5.py -k "not perm and not multi and not ibd"
timeout 900 python bench.py --steps 2 --warmup 2 --profile-host gpurun_out/host_profile.txt > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "bench_full rc=$?"; cat gpurun_out/bench_full.json; tail -5 gpurun_out/bench_full.err
export MMG_PROFILE_RANGE=1
timeout 1200 ncu --profile-from-start off --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:QuadEpi -c 1 \
   -o gpurun_out/prof_quad -f python bench.py --snps 131072 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/prof_quad.log 2>&1
echo "full capture quad rc=$?"; tail -3 gpurun_out/prof_quad.log
