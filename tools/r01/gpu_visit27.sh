#!/bin/bash
# linear pre-pass out of the scan epilogue + shared L2 prefetch rota: scan tests, role clocks, 1M bench for panel / pair
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo build failed; tail -20 gpurun_out/build.log; exit 1; }
timeout 600 python -m pytest tests/test_gpu_reml_scan.py tests/test_gpu_reference_pin.py tests/test_gpu_hdf5.py -x -q -m gpu -p no:cacheprovider --timeout 300 > gpurun_out/tests_scan.log 2>&1
echo "scan tests rc=$?"; tail -4 gpurun_out/tests_scan.log
for cfg in panel:1 panel:4 pair:1 pair:4; do
  sched=${cfg%%:*}; sh=${cfg##*:}
  MMG_SCAN_SCHED=$sched MMG_SCAN_PF_SHARE=$sh MMG_SCAN_DBG_CLOCKS=gpurun_out/clocks3_${sched}_$sh.txt timeout 200 python bench.py --snps 131072 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/bench_small_${sched}_$sh.json 2> gpurun_out/bench_small_${sched}_$sh.err
  python tools/summ_clocks.py gpurun_out/clocks3_${sched}_$sh.txt
  MMG_SCAN_SCHED=$sched MMG_SCAN_PF_SHARE=$sh timeout 300 python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-e2e > gpurun_out/bench_${sched}_$sh.json 2> gpurun_out/bench_${sched}_$sh.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_${sched}_$sh.json'))
    print('$sched pf_share=$sh: value %.0f ms/step %.1f scan_kernel %.2f S=%d stages %s'%(d['value'], d['ms_per_step'], d['roofline']['launch_ms'], d['roofline']['slices'], {k: round(1e3*v,1) for k,v in d['stage_seconds_per_step'].items() if v}))
except Exception as e:
    print('$sched $sh parse failed', e)
PY
done
timeout 300 python -m pytest tests/test_gpu_full_size.py -x -q -m gpu -p no:cacheprovider --timeout 280 > gpurun_out/tests_full.log 2>&1
echo "full-size rc=$?"; tail -3 gpurun_out/tests_full.log
