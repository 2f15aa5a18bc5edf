#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo build failed; tail -20 gpurun_out/build.log; exit 1; }
for v in "panel 8" "pair 8" "panel 6"; do
  set -- $v
  MMG_SCAN_SCHED=$1 MMG_SCAN_PANEL=$2 timeout 300 python -m pytest tests/test_gpu_reml_scan.py -x -q -m gpu -k "tcgen05 or agree or multi or perm" -p no:cacheprovider --timeout 200 > gpurun_out/tests_$1_$2.log 2>&1
  echo "tests $1 $2 rc=$?"; tail -2 gpurun_out/tests_$1_$2.log
done
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
bench() { name=$1; m=$2; shift; shift
  env "$@" timeout 300 python bench.py --snps $m --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_$name.json'))
    r=d['roofline']
    print('$name: value %.0f scan_ms %.2f S=%s gram_ms %.2f scan_stage %.1f clocks %s %s W'%(d['value'], r['launch_ms'], r['slices'], d['kinship']['gram_ms'], 1e3*d['stage_seconds_per_step']['scan'], d['clocks']['sm_mhz'], d['clocks']['power_w_max']))
except Exception as e: print('$name parse fail', e)
PY
}
clk panel8_pf8
clk panel8_pf0 MMG_SCAN_PREFETCH=0
clk pair8_pf0 MMG_SCAN_SCHED=pair MMG_SCAN_PREFETCH=0
bench panel8_pf8 262144 MMG_TC_SLICES=5
bench panel8_pf0 262144 MMG_TC_SLICES=5 MMG_SCAN_PREFETCH=0
bench pair8_pf0 262144 MMG_TC_SLICES=5 MMG_SCAN_PREFETCH=0 MMG_SCAN_SCHED=pair
bench full_pf0 1000000 MMG_SCAN_PREFETCH=0
bench full_pf8 1000000
