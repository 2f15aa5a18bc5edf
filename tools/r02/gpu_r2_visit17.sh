#!/bin/bash
# int8 issue rate with the scan's operand data (full-range digit bytes in B): peaks file + a short bench line carrying it.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo build failed; tail -20 gpurun_out/build.log; exit 1; }
timeout 300 python tools/measure_peaks.py > gpurun_out/peaks.log 2>&1; echo "peaks rc=$?"; grep -E "tops|sm_mhz|power|cap" gpurun_out/peaks.log
python - <<'PY'
import sys, time, subprocess, threading
sys.path.insert(0, '.')
from mixmogam_b200 import get_context
ctx = get_context(0)
def smi():
    return subprocess.run(['nvidia-smi','-i','0','--query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap','--format=csv,noheader,nounits'],capture_output=True,text=True).stdout.strip()
for which in ('imma_pair_sustained3000', 'imma_pair_digits_sustained3000', 'imma_tcgen05_digits_sustained3000'):
    samples=[]; stop=[]
    th=threading.Thread(target=lambda: [samples.append(smi()) or time.sleep(0.2) for _ in iter(lambda: bool(stop), True)], daemon=True); th.start()
    v = ctx.microbench(which)
    stop.append(1); th.join()
    print(which, '%.0f TOP/s' % v, 'smi under load:', samples[len(samples)//2:][:3])
PY
timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_digits.json 2> gpurun_out/bench_digits.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_digits.json').read().strip().splitlines()[-1])
r=d['roofline']
print('value %.0f ms/step %.1f scan %.1f ms achieved %.0f peak %.0f frac %.3f digits-peak %s frac %s' % (d['value'], d['ms_per_step'], r['launch_ms'], r['achieved'], r['peak'], r['frac'], r.get('int8_sustained_tops_digit_operands'), r.get('frac_of_rate_with_digit_operands')), d['clocks'])
PY
