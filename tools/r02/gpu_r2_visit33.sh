#!/bin/bash
# f_sf series branch up to q = 5: scan tests, the phenotype batch's statistics kernel, the bench step.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo build failed; tail -20 gpurun_out/build.log; exit 1; }
timeout 900 python -m pytest tests/test_gpu_reml_scan.py tests/test_gpu_emma.py tests/test_gpu_reference_pin.py -q -m gpu -p no:cacheprovider > gpurun_out/t_scan.log 2>&1; echo "t_scan rc=$?"; tail -3 gpurun_out/t_scan.log
MMG_SHARED_DEBUG=1 timeout 900 python tools/bench_multi.py --indivs 10000 --snps 1000000 --phenotypes 199 --single 1 --unshared 0 > gpurun_out/r02_multi_1m.json 2> gpurun_out/r02_multi_1m.err; echo "multi rc=$?"; grep "shared scan" gpurun_out/r02_multi_1m.err | tail -2; python -c "
import json; d=json.loads(open('gpurun_out/r02_multi_1m.json').read().strip().splitlines()[-1]); print({k: d[k] for k in ('emmax_multi_s','scan_kernels_ms','snp_tests_per_s_scan_stage','scan_cost_vs_one_single_scan','max_rel_err_neglog10p_vs_single')}, d['stage_seconds']['scan'])"
timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_fsf.json 2> gpurun_out/bench_fsf.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_fsf.json').read().strip().splitlines()[-1])
print('value %.0f ms/step %.1f scan kernel %.1f frac %.3f' % (d['value'], d['ms_per_step'], d['roofline']['launch_ms'], d['roofline']['frac']), d['clocks']['sm_mhz'])
PY
