#!/bin/bash
# epilogue with 32-column tcgen05.ld (MMG_SCAN_LD=32) x {panel, pair}: role clocks at m=131072, then the 1M bench; full-size test; whole suite
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo build failed; tail -20 gpurun_out/build.log; exit 1; }
MMG_SCAN_LD=32 timeout 300 python -m pytest tests/test_gpu_reml_scan.py -x -q -m gpu -k "schedules_agree or certified or golden" -p no:cacheprovider --timeout 200 > gpurun_out/tests_ld32.log 2>&1
echo "ld32 tests rc=$?"; tail -3 gpurun_out/tests_ld32.log
for cfg in panel:16 panel:32 pair:16 pair:32; do
  sched=${cfg%%:*}; ld=${cfg##*:}
  MMG_SCAN_SCHED=$sched MMG_SCAN_LD=$ld MMG_SCAN_DBG_CLOCKS=gpurun_out/clocks_${sched}_$ld.txt timeout 200 python bench.py --snps 131072 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/bench_small_${sched}_$ld.json 2> gpurun_out/bench_small_${sched}_$ld.err
  python tools/summ_clocks.py gpurun_out/clocks_${sched}_$ld.txt
  MMG_SCAN_SCHED=$sched MMG_SCAN_LD=$ld timeout 300 python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-e2e > gpurun_out/bench_${sched}_$ld.json 2> gpurun_out/bench_${sched}_$ld.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_${sched}_$ld.json'))
    print('$sched ld$ld: value %.0f ms/step %.1f scan_kernel %.2f gram %.2f stages %s'%(d['value'], d['ms_per_step'], d['roofline']['launch_ms'], d['kinship']['gram_ms'], {k: round(1e3*v,1) for k,v in d['stage_seconds_per_step'].items() if v}))
except Exception as e:
    print('$sched ld$ld parse failed', e)
PY
done
timeout 300 python -m pytest tests/test_gpu_full_size.py -x -q -m gpu -p no:cacheprovider --timeout 280 > gpurun_out/tests_full.log 2>&1
echo "full-size rc=$?"; tail -5 gpurun_out/tests_full.log
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider --timeout 600 --deselect tests/test_gpu_full_size.py > gpurun_out/tests_gpu.log 2>&1
echo "tests rc=$?"; tail -5 gpurun_out/tests_gpu.log
