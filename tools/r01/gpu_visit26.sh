#!/bin/bash
# finer epilogue clocks (x loads | FP64 x.v pass | accumulator drain) for panel and pair
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo build failed; tail -20 gpurun_out/build.log; exit 1; }
for cfg in panel:16 pair:16; do
  sched=${cfg%%:*}; ld=${cfg##*:}
  MMG_SCAN_SCHED=$sched MMG_SCAN_LD=$ld MMG_SCAN_DBG_CLOCKS=gpurun_out/clocks2_${sched}_$ld.txt timeout 200 python bench.py --snps 131072 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/bench_small_${sched}_$ld.json 2> gpurun_out/bench_small_${sched}_$ld.err
  python tools/summ_clocks.py gpurun_out/clocks2_${sched}_$ld.txt
  python -c "
import json; d=json.load(open('gpurun_out/bench_small_${sched}_$ld.json')); print('$sched scan_kernel %.2f ms slices %d'%(d['roofline']['launch_ms'], d['roofline']['slices']))"
done
