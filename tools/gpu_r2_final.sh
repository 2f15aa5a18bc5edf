#!/bin/bash
# Round-2 final state on one B200 at HEAD: the whole GPU suite including the full-size parity test, smoke, the default bench
# (e2e + variants + CPU arm), the reference arm, launch list of one step, ncu --set full of the Gram and the scan.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo build failed; tail -20 gpurun_out/build.log; exit 1; }
timeout 2400 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/t_all.log 2>&1; echo "t_all rc=$?"; tail -6 gpurun_out/t_all.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_n1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n1.json').read().strip().splitlines()[-1])
print('N=1 value %.0f ms/step %.1f scan_kernel %.1f frac %.3f e2e %.0f (%.1f ms)'%(d['value'], d['ms_per_step'], d['roofline']['launch_ms'], d['roofline']['frac'], d['e2e']['value'], d['e2e']['ms_per_step']))
print('kinship', d['kinship'])
print('stages', {k: round(1e3*v,2) for k,v in d['stage_seconds_per_step'].items() if v})
print('e2e stages', {k: round(1e3*v,2) for k,v in d['e2e']['stage_seconds_per_step'].items() if v})
for k,v in d['e2e'].get('other_host_buffers',{}).items(): print(' e2e', k, '%.0f /s %.1f ms' % (v['value'], v['ms_per_step']), {a: round(1e3*b,1) for a,b in v['stage_seconds_per_step'].items() if b})
print('cpu', d.get('cpu_baseline',{}) and d['cpu_baseline']['value'], 'clocks', d['clocks'])
PY
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; tail -c 600 gpurun_out/bench_ref.json
export MMG_PROFILE_RANGE=1 MMG_SCAN_COOP=0
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_1m.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/launches_bench.log 2>&1; echo "launch list rc=$?"
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"gram_pair_kernel" -s 4 -c 1 -f -o gpurun_out/prof_gram_1m python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/prof_gram_1m.log 2>&1; echo "ncu gram rc=$?"
ncu -i gpurun_out/prof_gram_1m.ncu-rep --page raw --csv > gpurun_out/prof_gram_1m_raw.csv 2>/dev/null
rm -f gpurun_out/prof_gram_1m.ncu-rep
