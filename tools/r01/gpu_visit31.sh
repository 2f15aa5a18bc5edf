#!/bin/bash
# 2-GPU bench as the driver launches it (pair scan, pre-pass, streamed two-lane Gram upload in the e2e leg)
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo build failed; tail -20 gpurun_out/build.log; exit 1; }
N=2
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
   bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "bench N=$N rc=$?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_n2.json'))
print('N=2 value %.0f ms/step %.1f scan_kernel %.1f stages %s e2e %.0f'%(d['value'], d['ms_per_step'], d['roofline']['launch_ms'], {k: round(1e3*v,1) for k,v in d['stage_seconds_per_step'].items() if v}, d['e2e']['value']))
PY
tail -3 gpurun_out/bench_n$N.err
timeout 300 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; cut -c1-400 gpurun_out/bench_ref.json
