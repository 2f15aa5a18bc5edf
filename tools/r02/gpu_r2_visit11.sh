#!/bin/bash
# Pack kernels underneath the Gram (side stream, two operand slots) and the L2 prefetch warp of the GEMM core: kinship tests,
# A/B of the two switches on the bench step, launch list with the scan included, the 1M-SNP phenotype batch with its
# sub-timers, pipe-rate peaks with the mxf4 issue rate.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo build failed; tail -20 gpurun_out/build.log; exit 1; }
timeout 900 python -m pytest tests/test_gpu_kinship.py tests/test_gpu_reference_pin.py -q -m gpu -p no:cacheprovider -x > gpurun_out/t_kin.log 2>&1; echo "t_kin rc=$?"; tail -5 gpurun_out/t_kin.log
run() {  # name, env...
  name=$1; shift
  env "$@" timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err; echo "bench $name rc=$?"
  python - <<PY
import json
d=json.loads(open('gpurun_out/bench_$name.json').read().strip().splitlines()[-1])
print('$name', 'value %.0f ms/step %.1f gram %.2f frac %.3f' % (d['value'], d['ms_per_step'], d['kinship']['gram_ms'], d['kinship']['frac']), {k: round(1e3*v,2) for k,v in d['stage_seconds_per_step'].items() if v}, d['clocks']['sm_mhz'])
PY
}
run ov1_pf8 MMG_GRAM_OVERLAP=1 MMG_GRAM_PREFETCH=8
run ov0_pf0 MMG_GRAM_OVERLAP=0 MMG_GRAM_PREFETCH=0
run ov1_pf0 MMG_GRAM_OVERLAP=1 MMG_GRAM_PREFETCH=0
run ov0_pf8 MMG_GRAM_OVERLAP=0 MMG_GRAM_PREFETCH=8
run ov1_pf24 MMG_GRAM_OVERLAP=1 MMG_GRAM_PREFETCH=24
MMG_PROFILE_RANGE=1 MMG_SCAN_COOP=0 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_1m.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/launches_bench.log 2>&1; echo "launch list rc=$?"
MMG_SHARED_DEBUG=1 timeout 900 python tools/bench_multi.py --indivs 10000 --snps 1000000 --phenotypes 199 --single 1 --unshared 0 > gpurun_out/r02_multi_1m.json 2> gpurun_out/r02_multi_1m.err; echo "multi rc=$?"; tail -c 1200 gpurun_out/r02_multi_1m.json; grep "shared scan" gpurun_out/r02_multi_1m.err | tail -4
timeout 300 python tools/measure_peaks.py > gpurun_out/peaks.log 2>&1; echo "peaks rc=$?"; grep -E "mxf4|int8_sust|sm_mhz_median|power" gpurun_out/peaks.log
