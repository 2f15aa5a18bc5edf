#!/bin/bash
# 2-GPU bench at HEAD, as the driver launches it
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo build failed; tail -20 gpurun_out/build.log; exit 1; }
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
   bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
echo "bench N=2 rc=$? stdout lines: $(wc -l < gpurun_out/bench_n2.json)"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n2.json').read())
print('N=2 value %.0f ms/step %.1f scan_kernel %.1f e2e %.0f (%.1f ms) stages %s'%(d['value'], d['ms_per_step'], d['roofline']['launch_ms'], d['e2e']['value'], d['e2e']['ms_per_step'], {k: round(1e3*v,1) for k,v in d['stage_seconds_per_step'].items() if v}))
PY
