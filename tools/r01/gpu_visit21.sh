#!/bin/bash
# sharded R'R preparation: single-GPU equivalence test, then the 2-GPU bench as the driver launches it
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo build failed; tail -20 gpurun_out/build.log; exit 1; }
timeout 300 python -m pytest tests/test_gpu_reml_scan.py -x -q -m gpu -k "quad_form or golden" -p no:cacheprovider --timeout 200 > gpurun_out/tests_quadform.log 2>&1
echo "tests rc=$?"; tail -3 gpurun_out/tests_quadform.log
N=2
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
   bench.py --gpus $N --steps 2 --warmup 1 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "bench N=$N rc=$?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_n2.json'))
print('N=2 value %.0f ms/step %.1f scan_kernel %.1f stages %s e2e %.0f'%(d['value'], d['ms_per_step'], d['roofline']['launch_ms'], {k: round(1e3*v,1) for k,v in d['stage_seconds_per_step'].items()}, d['e2e']['value']))
PY
tail -3 gpurun_out/bench_n$N.err
