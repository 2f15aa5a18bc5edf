#!/bin/bash
# Statistics deferred out of the scan's epilogue (quad_finish_kernel): scan tests, A/B on the bench step (same box, alternating).
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo build failed; tail -20 gpurun_out/build.log; exit 1; }
timeout 900 python -m pytest tests/test_gpu_reml_scan.py tests/test_gpu_emma.py tests/test_gpu_reference_pin.py -q -m gpu -p no:cacheprovider > gpurun_out/t_scan.log 2>&1; echo "t_scan rc=$?"; tail -3 gpurun_out/t_scan.log
for d in 1 0 1 0; do
  MMG_SCAN_DEFER_STATS=$d timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_defer$d.json 2> gpurun_out/bench_defer$d.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/bench_defer$d.json').read().strip().splitlines()[-1])
print('defer=$d value %.0f ms/step %.1f scan kernel %.1f frac %.3f clocks %s' % (d['value'], d['ms_per_step'], d['roofline']['launch_ms'], d['roofline']['frac'], d['clocks']['sm_mhz']))
PY
done
