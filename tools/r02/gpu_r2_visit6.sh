#!/bin/bash
# Round 2, visit 6: shared-rotation phenotype batch -- tests, then configs[2] at n = 10 000 (T = 199) on a SNP subset.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo build failed; tail -20 gpurun_out/build.log; exit 1; }
timeout 900 python -m pytest tests/test_gpu_emma.py tests/test_gpu_reml_scan.py -q -m gpu -p no:cacheprovider -k "two_env or multi" > gpurun_out/t_new.log 2>&1; echo "t_new rc=$?"; tail -12 gpurun_out/t_new.log
timeout 900 python tools/bench_multi.py ${MULTI_ARGS:---indivs 10000 --snps 131072 --phenotypes 199} > gpurun_out/r02_multi.json 2> gpurun_out/r02_multi.err; echo "multi rc=$?"; tail -c 2500 gpurun_out/r02_multi.json; tail -5 gpurun_out/r02_multi.err
