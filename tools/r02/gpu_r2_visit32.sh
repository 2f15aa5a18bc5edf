#!/bin/bash
# ncu --set full of the phenotype batch's statistics kernel (shared_finish_kernel), n = 10k, T = 199, 131k SNPs.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo build failed; tail -20 gpurun_out/build.log; exit 1; }
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"shared_finish_kernel" -s 2 -c 1 -f -o gpurun_out/prof_finish python tools/bench_multi.py --indivs 10000 --snps 131072 --phenotypes 199 --single 0 --unshared 0 > gpurun_out/prof_finish.log 2>&1; echo "ncu rc=$?"
ncu -i gpurun_out/prof_finish.ncu-rep --page raw --csv > gpurun_out/prof_finish_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_finish.ncu-rep --page details > gpurun_out/prof_finish_details.txt 2>/dev/null
ncu -i gpurun_out/prof_finish.ncu-rep --page source --csv > gpurun_out/prof_finish_source.csv 2>/dev/null
python tools/ncu_summary.py raw gpurun_out/prof_finish_raw.csv
grep -E "Duration|Registers|Theoretical Occ|Achieved Occ|L1/TEX Hit|Executed Ipc|Local" gpurun_out/prof_finish_details.txt | head -12
rm -f gpurun_out/prof_finish.ncu-rep
