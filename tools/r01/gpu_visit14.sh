#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo build failed; tail -20 gpurun_out/build.log; exit 1; }
timeout 600 python -m pytest tests/test_gpu_reml_scan.py tests/test_gpu_reference_pin.py -x -q -m gpu -p no:cacheprovider --timeout 200 > gpurun_out/tests_scan.log 2>&1
echo "tests rc=$?"; tail -2 gpurun_out/tests_scan.log
bench() { name=$1; shift
  env "$@" timeout 600 python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-e2e > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_$name.json'))
    r=d['roofline']
    print('$name: value %.0f scan_ms %.2f frac %.3f S=%s gram_ms %.2f scan_stage %.1f clocks %s %s W'%(d['value'], r['launch_ms'], r['frac'], r['slices'], d['kinship']['gram_ms'], 1e3*d['stage_seconds_per_step']['scan'], d['clocks']['sm_mhz'], d['clocks']['power_w_max']))
except Exception as e: print('$name parse fail', e)
PY
}
bench sync1 MMG_SCAN_WAVE_SYNC=1
bench sync0 MMG_SCAN_WAVE_SYNC=0
bench sync1_pf32 MMG_SCAN_WAVE_SYNC=1 MMG_SCAN_PREFETCH=32
bench sync1_again MMG_SCAN_WAVE_SYNC=1
export MMG_PROFILE_RANGE=1
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__cycles_elapsed.avg.per_second --clock-control none -k regex:scan_quad_kernel -c 2 --csv \
   --log-file gpurun_out/scan_sync1_metrics.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > /dev/null 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/scan_sync1_metrics.csv | cut -c1-600
