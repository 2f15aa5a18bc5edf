#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
run() { name=$1; shift; timeout 900 python -m pytest "$@" -q -m gpu -p no:cacheprovider --timeout 600 > gpurun_out/$name.log 2>&1; echo "$name rc=$?"; tail -25 gpurun_out/$name.log; }
run t_pin tests/test_gpu_reference_pin.py
run t_multi tests/test_gpu_reml_scan.py -k "multi"
run t_kin tests/test_gpu_kinship.py -k "resident or golden"
for h in "normal normal" "last first" "last normal" "normal first" "first last"; do
  set -- $h
  MMG_TC_HINT_A=$1 MMG_TC_HINT_B=$2 timeout 600 python bench.py --snps 262144 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/bench_hint_$1_$2.json 2> gpurun_out/bench_hint.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_hint_$1_$2.json'))
    print('hint A=$1 B=$2 value %.0f scan_ms %.2f frac %.3f gram_ms %.2f clocks %s'%(d['value'], d['roofline']['launch_ms'], d['roofline']['frac'], d['kinship']['gram_ms'], d['clocks']))
except Exception as e: print('parse fail', e)
PY
done
MMG_H2D=2d timeout 900 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_h2d2d.json 2> gpurun_out/bench_h2d2d.err; echo "bench 2d rc=$?"; cat gpurun_out/bench_h2d2d.json
timeout 900 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "bench_full rc=$?"; cat gpurun_out/bench_full.json; tail -5 gpurun_out/bench_full.err
