#!/bin/bash
# e2m1 / kind::mxf4 Gram: bit-exact tests, A/B against the int8 Gram on the full bench step, launch list of one timed step,
# ncu --set full of the scan (cooperative launch off: ncu's replay cannot co-schedule the grid) and of the mxf4 Gram.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo build failed; tail -20 gpurun_out/build.log; exit 1; }
timeout 900 python -m pytest tests/test_gpu_kinship.py -q -m gpu -p no:cacheprovider -x > gpurun_out/t_kin.log 2>&1; echo "t_kin rc=$?"; tail -15 gpurun_out/t_kin.log
for kind in fp4 i8; do
  MMG_GRAM_KIND=$kind timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_gram_$kind.json 2> gpurun_out/bench_gram_$kind.err; echo "bench $kind rc=$?"
  python - <<PY
import json
d=json.loads(open('gpurun_out/bench_gram_$kind.json').read().strip().splitlines()[-1])
print('$kind', 'value %.0f ms/step %.1f' % (d['value'], d['ms_per_step']), d['kinship'], {k: round(1e3*v,2) for k,v in d['stage_seconds_per_step'].items() if v}, d['clocks'])
PY
done
export MMG_PROFILE_RANGE=1
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_1m.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/launches_bench.log 2>&1; echo "launch list rc=$?"
MMG_SCAN_COOP=0 timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"scan_quad_kernel" -c 1 -f -o gpurun_out/prof_scan_1m python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/prof_scan_1m.log 2>&1; echo "ncu scan rc=$?"
ncu -i gpurun_out/prof_scan_1m.ncu-rep --page raw --csv > gpurun_out/prof_scan_1m_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_scan_1m.ncu-rep --page details > gpurun_out/prof_scan_1m_details.txt 2>/dev/null
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"tc_gemm_i8_kernel" -s 4 -c 1 -f -o gpurun_out/prof_gram_1m python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/prof_gram_1m.log 2>&1; echo "ncu gram rc=$?"
ncu -i gpurun_out/prof_gram_1m.ncu-rep --page raw --csv > gpurun_out/prof_gram_1m_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_gram_1m.ncu-rep --page details > gpurun_out/prof_gram_1m_details.txt 2>/dev/null
