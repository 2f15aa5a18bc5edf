#!/bin/bash
# CTA-pair scan kernel: parity, sweep, ncu full capture of panel and pair kernels
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo build failed; tail -20 gpurun_out/build.log; exit 1; }
for p in 8 6 4; do
  MMG_SCAN_SCHED=pair MMG_SCAN_PANEL=$p timeout 300 python -m pytest tests/test_gpu_reml_scan.py -x -q -m gpu -k "tcgen05 or agree or multi or perm" -p no:cacheprovider --timeout 120 > gpurun_out/tests_pair_p$p.log 2>&1
  echo "tests pair panel=$p rc=$?"; tail -3 gpurun_out/tests_pair_p$p.log
done
bench() { name=$1; shift
  env "$@" timeout 300 python bench.py --snps 262144 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_$name.json'))
    print('$name: value %.0f scan_ms %.2f frac %.3f gram_ms %.2f clocks %s'%(d['value'], d['roofline']['launch_ms'], d['roofline']['frac'], d['kinship']['gram_ms'], d['clocks']))
except Exception as e: print('$name parse fail', e)
PY
}
bench panel8 MMG_SCAN_PANEL=8
bench pair8 MMG_SCAN_SCHED=pair MMG_SCAN_PANEL=8
bench pair6 MMG_SCAN_SCHED=pair MMG_SCAN_PANEL=6
bench pair4 MMG_SCAN_SCHED=pair MMG_SCAN_PANEL=4
bench pair8_nopf MMG_SCAN_SCHED=pair MMG_SCAN_PANEL=8 MMG_SCAN_PREFETCH=0
export MMG_PROFILE_RANGE=1
for v in "panel 8" "pair 8"; do
  set -- $v
  MMG_SCAN_SCHED=$1 MMG_SCAN_PANEL=$2 timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:scan_quad_kernel -c 1 \
     -o gpurun_out/prof_scan_$1 -f python bench.py --snps 65536 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/prof_scan_$1.log 2>&1
  echo "ncu $1 rc=$?"
done
ls -la gpurun_out | tail -8
