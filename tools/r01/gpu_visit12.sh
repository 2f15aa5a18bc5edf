#!/bin/bash
# diagonal split + certified plane count + genotype prefetch: parity everywhere, then sweep and the full bench
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo build failed; tail -20 gpurun_out/build.log; exit 1; }
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider --timeout 300 -x > gpurun_out/tests_all.log 2>&1
echo "tests all rc=$?"; tail -3 gpurun_out/tests_all.log
for v in "pair 8" "table 8" "panel 6"; do
  set -- $v
  MMG_SCAN_SCHED=$1 MMG_SCAN_PANEL=$2 timeout 300 python -m pytest tests/test_gpu_reml_scan.py -x -q -m gpu -k "tcgen05 or agree or multi or perm" -p no:cacheprovider --timeout 200 > gpurun_out/tests_$1_$2.log 2>&1
  echo "tests $1 $2 rc=$?"; tail -2 gpurun_out/tests_$1_$2.log
done
bench() { name=$1; shift
  env "$@" timeout 300 python bench.py --snps 262144 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_$name.json'))
    r=d['roofline']
    print('$name: value %.0f scan_ms %.2f frac %.3f S=%s rho=%.2e gram_ms %.2f scan_stage %.1f clocks %s'%(d['value'], r['launch_ms'], r['frac'], r['slices'], r['certified_rel_bound_xx'], d['kinship']['gram_ms'], 1e3*d['stage_seconds_per_step']['scan'], d['clocks']['sm_mhz']))
except Exception as e: print('$name parse fail', e)
PY
}
bench auto
bench S7 MMG_TC_SLICES=7
bench S6 MMG_TC_SLICES=6
bench auto_c4 MMG_SCAN_CLUSTER=4
bench auto_tol9 MMG_TC_TOL=1e-9
timeout 900 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "full rc=$?"; cat gpurun_out/bench_full.json
