#!/bin/bash
# R'R on the int8 tensor pipe: new test, scan suite, full-size test, bench (int8 vs dsyrk)
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo build failed; tail -20 gpurun_out/build.log; exit 1; }
timeout 300 python -m pytest tests/test_gpu_reml_scan.py -x -q -m gpu -k "quad_form" -p no:cacheprovider --timeout 200 > gpurun_out/tests_quad.log 2>&1
echo "quad tests rc=$?"; tail -12 gpurun_out/tests_quad.log
timeout 600 python -m pytest tests/test_gpu_reml_scan.py tests/test_gpu_reference_pin.py tests/test_gpu_hdf5.py tests/test_gpu_full_size.py -q -m gpu -p no:cacheprovider --timeout 300 > gpurun_out/tests_scan.log 2>&1
echo "scan tests rc=$?"; tail -5 gpurun_out/tests_scan.log
for a in int8 dsyrk; do
MMG_QUAD_A=$a timeout 300 python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-e2e > gpurun_out/bench_a_$a.json 2> gpurun_out/bench_a_$a.err
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_a_$a.json'))
    print('A=$a: value %.0f ms/step %.1f scan_kernel %.2f S=%d rho %.3g stages %s'%(d['value'], d['ms_per_step'], d['roofline']['launch_ms'], d['roofline']['slices'], d['roofline']['certified_rel_bound_xx'], {k: round(1e3*v,1) for k,v in d['stage_seconds_per_step'].items() if v}))
except Exception as e:
    print('$a parse failed', e)
PY
tail -2 gpurun_out/bench_a_$a.err
done
