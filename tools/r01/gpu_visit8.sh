#!/bin/bash
# what bounds the scan kernel: MMA issue-rate microbenchmarks (with / without TMEM read-back), epilogue-reduced timing runs
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo build failed; tail -20 gpurun_out/build.log; exit 1; }
python - <<'PY' > gpurun_out/microbench_imma.txt 2>&1
import sys
sys.path.insert(0, '.')
from mixmogam_b200 import _lib
ctx = _lib.get_context(0)
for w in ['imma_tcgen05', 'imma_pair', 'imma_tcgen05_ldtm8', 'imma_pair_ldtm8', 'imma_tcgen05_ldtm4', 'imma_tcgen05_ldtm2', 'imma_tcgen05_ldtm16', 'imma_tcgen05', 'dmma', 'copy']:
    try:
        print(w, '%.1f' % ctx.microbench(w), flush=True)
    except Exception as e:
        print(w, 'FAILED', e, flush=True)
PY
cat gpurun_out/microbench_imma.txt
MMG_SCAN_SCHED=pair MMG_SCAN_PANEL=8 timeout 300 python -m pytest tests/test_gpu_reml_scan.py -x -q -m gpu -k "tcgen05 or agree or multi or perm" -p no:cacheprovider --timeout 200 > gpurun_out/tests_pair_p8.log 2>&1
echo "tests pair panel=8 rc=$?"; tail -3 gpurun_out/tests_pair_p8.log
bench() { name=$1; shift
  env "$@" timeout 300 python bench.py --snps 131072 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_$name.json'))
    print('$name: value %.0f scan_ms %.2f frac %.3f gram_ms %.2f clocks %s'%(d['value'], d['roofline']['launch_ms'], d['roofline']['frac'], d['kinship']['gram_ms'], d['clocks']))
except Exception as e: print('$name parse fail', e)
PY
}
bench panel8 MMG_SCAN_PANEL=8
bench panel8_epi1 MMG_SCAN_PANEL=8 MMG_SCAN_DBG_EPI=1
bench panel8_epi0 MMG_SCAN_PANEL=8 MMG_SCAN_DBG_EPI=0
bench panel8_epi4 MMG_SCAN_PANEL=8 MMG_SCAN_DBG_EPI=4
bench pair8_epi0 MMG_SCAN_SCHED=pair MMG_SCAN_PANEL=8 MMG_SCAN_DBG_EPI=0
bench table_cs2 MMG_SCAN_SCHED=table
bench table_cs4 MMG_SCAN_SCHED=table MMG_SCAN_CLUSTER=4
