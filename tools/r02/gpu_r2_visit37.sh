#!/bin/bash
# Default bench line at HEAD (round-2 final) + reference arm.
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
print('cpu', d.get('cpu_baseline',{}) and d['cpu_baseline']['value'], 'clocks', d['clocks'], 'traffic', d['roofline']['traffic'])
PY
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/bench_ref.json').read().strip().splitlines()[-1]); print(d['value'], d['cpu_baseline']['cores'])"
