#!/bin/bash
# Full-size parity test at HEAD (n = 10 000 x 1M SNPs) and the ncu --set full capture of the CTA-pair e2m1 Gram.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo build failed; tail -20 gpurun_out/build.log; exit 1; }
timeout 1200 python -m pytest tests/test_gpu_full_size.py -q -m gpu -p no:cacheprovider > gpurun_out/t_full.log 2>&1; echo "t_full rc=$?"; tail -4 gpurun_out/t_full.log
export MMG_PROFILE_RANGE=1 MMG_SCAN_COOP=0
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"gram_pair_kernel" -s 4 -c 1 -f -o gpurun_out/prof_gram_1m python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/prof_gram_1m.log 2>&1; echo "ncu gram rc=$?"
ncu -i gpurun_out/prof_gram_1m.ncu-rep --page raw --csv > gpurun_out/prof_gram_1m_raw.csv 2>/dev/null
rm -f gpurun_out/prof_gram_1m.ncu-rep
