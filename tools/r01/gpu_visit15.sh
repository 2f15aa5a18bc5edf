#!/bin/bash
# 128-column tile variants (4 accumulator stages): parity + sweep at m=262144 (S fixed at 5 for comparability) + 1M for the best
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo build failed; tail -20 gpurun_out/build.log; exit 1; }
for v in "pair128 8" "pair128 10" "pair128 12" "n128 8" "panel 8"; do
  set -- $v
  MMG_SCAN_SCHED=$1 MMG_SCAN_PANEL=$2 timeout 300 python -m pytest tests/test_gpu_reml_scan.py -x -q -m gpu -k "tcgen05 or agree or multi or perm" -p no:cacheprovider --timeout 200 > gpurun_out/tests_$1_$2.log 2>&1
  echo "tests $1 $2 rc=$?"; tail -2 gpurun_out/tests_$1_$2.log
done
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
bench panel8 262144 MMG_TC_SLICES=5
bench pair8 262144 MMG_TC_SLICES=5 MMG_SCAN_SCHED=pair
bench pair128_8 262144 MMG_TC_SLICES=5 MMG_SCAN_SCHED=pair128 MMG_SCAN_PANEL=8
bench pair128_10 262144 MMG_TC_SLICES=5 MMG_SCAN_SCHED=pair128 MMG_SCAN_PANEL=10
bench pair128_12 262144 MMG_TC_SLICES=5 MMG_SCAN_SCHED=pair128 MMG_SCAN_PANEL=12
bench pair128_6 262144 MMG_TC_SLICES=5 MMG_SCAN_SCHED=pair128 MMG_SCAN_PANEL=6
bench n128_8 262144 MMG_TC_SLICES=5 MMG_SCAN_SCHED=n128 MMG_SCAN_PANEL=8
bench n128_10 262144 MMG_TC_SLICES=5 MMG_SCAN_SCHED=n128 MMG_SCAN_PANEL=10
bench full_pair128_10 1000000 MMG_SCAN_SCHED=pair128 MMG_SCAN_PANEL=10
bench full_panel8 1000000
