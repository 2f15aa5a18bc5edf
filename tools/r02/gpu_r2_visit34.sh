#!/bin/bash
# After the f_sf change: default bench line (e2e, variants, CPU arm), launch list, ncu --set full of the scan kernel.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo build failed; tail -20 gpurun_out/build.log; exit 1; }
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_n1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n1.json').read().strip().splitlines()[-1])
print('N=1 value %.0f ms/step %.1f scan_kernel %.1f frac %.3f (digits %.3f) e2e %.0f (%.1f ms)'%(d['value'], d['ms_per_step'], d['roofline']['launch_ms'], d['roofline']['frac'], d['roofline']['frac_of_rate_with_digit_operands'], d['e2e']['value'], d['e2e']['ms_per_step']))
print('stages', {k: round(1e3*v,2) for k,v in d['stage_seconds_per_step'].items() if v})
print('e2e stages', {k: round(1e3*v,2) for k,v in d['e2e']['stage_seconds_per_step'].items() if v})
for k,v in d['e2e'].get('other_host_buffers',{}).items(): print(' e2e', k, '%.0f /s %.1f ms' % (v['value'], v['ms_per_step']))
print('cpu', d.get('cpu_baseline',{}) and d['cpu_baseline']['value'], 'clocks', d['clocks'])
PY
export MMG_PROFILE_RANGE=1 MMG_SCAN_COOP=0
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_1m.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/launches_bench.log 2>&1; echo "launch list rc=$?"
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"scan_quad_kernel" -c 1 -f -o gpurun_out/prof_scan_1m python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/prof_scan_1m.log 2>&1; echo "ncu scan rc=$?"
ncu -i gpurun_out/prof_scan_1m.ncu-rep --page raw --csv > gpurun_out/prof_scan_1m_raw.csv 2>/dev/null
rm -f gpurun_out/prof_scan_1m.ncu-rep
