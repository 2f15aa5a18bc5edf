#!/bin/bash
# pre-pass on a side stream underneath the int8 R'R: scan tests, bench with and without the overlap
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo build failed; tail -20 gpurun_out/build.log; exit 1; }
timeout 600 python -m pytest tests/test_gpu_reml_scan.py tests/test_gpu_reference_pin.py tests/test_gpu_hdf5.py tests/test_gpu_full_size.py -q -m gpu -p no:cacheprovider --timeout 300 > gpurun_out/tests_scan.log 2>&1
echo "scan tests rc=$?"; tail -4 gpurun_out/tests_scan.log
for ov in 1 0; do
MMG_SCAN_OVERLAP=$ov timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e > gpurun_out/bench_ov$ov.json 2> gpurun_out/bench_ov$ov.err
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_ov$ov.json'))
    print('overlap=$ov: value %.0f ms/step %.1f scan_kernel %.2f S=%d stages %s'%(d['value'], d['ms_per_step'], d['roofline']['launch_ms'], d['roofline']['slices'], {k: round(1e3*v,1) for k,v in d['stage_seconds_per_step'].items() if v}))
except Exception as e:
    print('$ov parse failed', e)
PY
tail -2 gpurun_out/bench_ov$ov.err
done
