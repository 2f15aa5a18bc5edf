#!/bin/bash
# per-role wait-cycle counters of the scan kernel
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo build failed; tail -20 gpurun_out/build.log; exit 1; }
bench() { name=$1; shift
  env "$@" MMG_SCAN_DBG_CLOCKS=gpurun_out/clocks_$name.txt timeout 300 python bench.py --snps 131072 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err
  python - <<PY
import json, numpy as np
try:
    d=json.load(open('gpurun_out/bench_$name.json'))
    print('$name: value %.0f scan_ms %.2f frac %.3f clocks %s'%(d['value'], d['roofline']['launch_ms'], d['roofline']['frac'], d['clocks']['sm_mhz']))
    a=np.loadtxt('gpurun_out/clocks_$name.txt')
    names='cta prod_total prod_wait_empty prod_wait_aempty - mma_total mma_wait_full mma_wait_tempty mma_wait_afull epi_total epi_wait_tfull'.split()
    ev=a[a[:,5]>0]
    print('   rows with mma', len(ev), ' '.join('%s=%.2fM'%(names[i], ev[:,i].mean()/1e6) for i in (1,2,3,5,6,7,8)), ' '.join('%s=%.2fM'%(names[i], a[:,i].mean()/1e6) for i in (9,10)))
except Exception as e: print('$name parse fail', e)
PY
}
bench panel8 MMG_SCAN_PANEL=8
bench pair8 MMG_SCAN_SCHED=pair MMG_SCAN_PANEL=8
bench pair8_epi0 MMG_SCAN_SCHED=pair MMG_SCAN_PANEL=8 MMG_SCAN_DBG_EPI=0
bench pair8_nopf MMG_SCAN_SCHED=pair MMG_SCAN_PANEL=8 MMG_SCAN_PREFETCH=0
bench pair6 MMG_SCAN_SCHED=pair MMG_SCAN_PANEL=6
bench pair8_S1 MMG_SCAN_SCHED=pair MMG_SCAN_PANEL=8 MMG_TC_SLICES=1
