#!/bin/bash
# genotype-stationary scan kernel: parity under every (panel, cluster) variant, then a schedule sweep at m=262144
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo build failed; tail -20 gpurun_out/build.log; exit 1; }
timeout 900 python -m pytest tests/test_gpu_reml_scan.py tests/test_gpu_reference_pin.py tests/test_gpu_hdf5.py -q -m gpu -p no:cacheprovider --timeout 300 > gpurun_out/tests_default.log 2>&1
echo "tests default rc=$?"; tail -4 gpurun_out/tests_default.log
for v in "8 2" "4 2" "6 1" "6 4" "8 4"; do
  set -- $v
  MMG_SCAN_PANEL=$1 MMG_SCAN_CLUSTER=$2 timeout 600 python -m pytest tests/test_gpu_reml_scan.py -x -q -m gpu -k "tcgen05 or agree or multi or perm" -p no:cacheprovider --timeout 300 > gpurun_out/tests_p$1_c$2.log 2>&1
  echo "tests panel=$1 cs=$2 rc=$?"; tail -2 gpurun_out/tests_p$1_c$2.log
done
bench() { name=$1; shift
  env "$@" timeout 600 python bench.py --snps 262144 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_$name.json'))
    print('$name: value %.0f scan_ms %.2f frac %.3f gram_ms %.2f clocks %s'%(d['value'], d['roofline']['launch_ms'], d['roofline']['frac'], d['kinship']['gram_ms'], d['clocks']))
except Exception as e: print('$name parse fail', e)
PY
}
bench table MMG_SCAN_SCHED=table
bench p6c2 MMG_SCAN_PANEL=6
bench p6c2_nopf MMG_SCAN_PANEL=6 MMG_SCAN_PREFETCH=0
bench p6c2_pf16 MMG_SCAN_PANEL=6 MMG_SCAN_PREFETCH=16
bench p8c2 MMG_SCAN_PANEL=8
bench p4c2 MMG_SCAN_PANEL=4
bench p6c4 MMG_SCAN_PANEL=6 MMG_SCAN_CLUSTER=4
bench p6c1 MMG_SCAN_PANEL=6 MMG_SCAN_CLUSTER=1
bench p6c2_Bnormal MMG_SCAN_PANEL=6 MMG_TC_HINT_B=normal
bench p6c2_S6 MMG_SCAN_PANEL=6 MMG_TC_SLICES=6
