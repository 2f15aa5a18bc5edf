#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo build failed; tail -20 gpurun_out/build.log; exit 1; }
clk() { name=$1; shift
env "$@" MMG_TC_SLICES=5 MMG_SCAN_DBG_CLOCKS=gpurun_out/clocks_$name.txt timeout 300 python bench.py --snps 131072 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/bench_${name}_clk.json 2>/dev/null
python - <<PY
import numpy as np, json
a=np.loadtxt('gpurun_out/clocks_$name.txt')
names='cta prod_total prod_wait_empty prod_wait_aempty - mma_total mma_wait_full mma_wait_tempty mma_wait_afull epi_total epi_wait_tfull'.split()
ev=a[a[:,5]>0]
d=json.load(open('gpurun_out/bench_${name}_clk.json'))
print('$name scan_ms %.2f:'%d['roofline']['launch_ms'], ' '.join('%s=%.2fM'%(names[i], ev[:,i].mean()/1e6) for i in (1,2,3,5,6,7,8)), ' '.join('%s=%.2fM'%(names[i], a[:,i].mean()/1e6) for i in (9,10)))
PY
}
clk panel8
clk pair8 MMG_SCAN_SCHED=pair
clk pair128_10 MMG_SCAN_SCHED=pair128 MMG_SCAN_PANEL=10
clk n128_8 MMG_SCAN_SCHED=n128
