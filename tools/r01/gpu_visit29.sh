#!/bin/bash
# dedicated L2-prefetch warp + FP64-only pre-pass: scan tests, role clocks, 1M bench sweeps (prefetch distance / rota), launch list
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo build failed; tail -20 gpurun_out/build.log; exit 1; }
timeout 600 python -m pytest tests/test_gpu_reml_scan.py tests/test_gpu_hdf5.py -x -q -m gpu -p no:cacheprovider --timeout 300 > gpurun_out/tests_scan.log 2>&1
echo "scan tests rc=$?"; tail -4 gpurun_out/tests_scan.log
for cfg in pair:8:1 panel:8:1; do
  IFS=: read sched pf sh <<< "$cfg"
  MMG_SCAN_SCHED=$sched MMG_SCAN_PREFETCH=$pf MMG_SCAN_PF_SHARE=$sh MMG_SCAN_DBG_CLOCKS=gpurun_out/clocks4_${sched}_${pf}_$sh.txt timeout 200 python bench.py --snps 131072 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/bench_small.json 2> gpurun_out/bench_small.err
  python tools/summ_clocks.py gpurun_out/clocks4_${sched}_${pf}_$sh.txt
done
for cfg in pair:8:1 pair:16:1 pair:32:1 pair:16:2 pair:0:1 panel:16:1; do
  IFS=: read sched pf sh <<< "$cfg"
  MMG_SCAN_SCHED=$sched MMG_SCAN_PREFETCH=$pf MMG_SCAN_PF_SHARE=$sh timeout 300 python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-e2e > gpurun_out/bench_${sched}_${pf}_$sh.json 2> gpurun_out/bench_${sched}_${pf}_$sh.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_${sched}_${pf}_$sh.json'))
    print('$sched prefetch=$pf share=$sh: value %.0f ms/step %.1f scan_kernel %.2f S=%d stages %s'%(d['value'], d['ms_per_step'], d['roofline']['launch_ms'], d['roofline']['slices'], {k: round(1e3*v,1) for k,v in d['stage_seconds_per_step'].items() if v}))
except Exception as e:
    print('$cfg parse failed', e)
PY
done
export MMG_PROFILE_RANGE=1
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
   --log-file gpurun_out/launches_1m.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/launches_bench.log 2>&1
echo "launch list rc=$?"; python tools/ncu_summary.py launches gpurun_out/launches_1m.csv | head -8
