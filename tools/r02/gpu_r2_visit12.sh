#!/bin/bash
# Pack kernels underneath the Gram (fixed side-stream set-up), AUTO -> FP64 scan for short scans: tests, A/B, launch list with the
# scan included, 1M-SNP phenotype batch with sub-timers.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo build failed; tail -20 gpurun_out/build.log; exit 1; }
timeout 900 python -m pytest tests/test_gpu_kinship.py tests/test_gpu_reference_pin.py tests/test_gpu_reml_scan.py -q -m gpu -p no:cacheprovider > gpurun_out/t_kin.log 2>&1; echo "t_kin rc=$?"; tail -8 gpurun_out/t_kin.log
run() {  # name, env...
  name=$1; shift
  env "$@" timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err; echo "bench $name rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_$name.json').read().strip().splitlines()[-1])
    print('$name', 'value %.0f ms/step %.1f gram %.2f frac %.3f' % (d['value'], d['ms_per_step'], d['kinship']['gram_ms'], d['kinship']['frac']), {k: round(1e3*v,2) for k,v in d['stage_seconds_per_step'].items() if v}, d['clocks']['sm_mhz'])
except Exception as e:
    print('$name failed', e)
PY
}
run ov1 MMG_GRAM_OVERLAP=1
run ov0 MMG_GRAM_OVERLAP=0
run ov1_b MMG_GRAM_OVERLAP=1
MMG_PROFILE_RANGE=1 MMG_SCAN_COOP=0 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_1m.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/launches_bench.log 2>&1; echo "launch list rc=$?"
MMG_SHARED_DEBUG=1 timeout 900 python tools/bench_multi.py --indivs 10000 --snps 1000000 --phenotypes 199 --single 1 --unshared 0 > gpurun_out/r02_multi_1m.json 2> gpurun_out/r02_multi_1m.err; echo "multi rc=$?"; tail -c 1200 gpurun_out/r02_multi_1m.json; grep "shared scan" gpurun_out/r02_multi_1m.err | tail -4
