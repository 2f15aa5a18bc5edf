#!/bin/bash
# N-GPU visit (gpurun --gpus N): the sharded-path test and bench.py under torchrun as the driver launches it.
N=${1:-2}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo build failed; tail -20 gpurun_out/build.log; exit 1; }
timeout 600 python -m pytest tests/test_gpu_multi.py -q -m gpu -p no:cacheprovider -x > gpurun_out/t_multi.log 2>&1; echo "t_multi rc=$?"; tail -15 gpurun_out/t_multi.log
for n in $(seq 2 2 $N | tr '\n' ' '); do
  if [ $n -eq 6 ]; then continue; fi
  MMG_BENCH_DEBUG=gpurun_out timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n \
     bench.py --gpus $n --steps ${STEPS:-10} --warmup 3 > gpurun_out/bench_n$n.json 2> gpurun_out/bench_n$n.err
  echo "bench N=$n rc=$? stdout lines: $(wc -l < gpurun_out/bench_n$n.json)"; tail -5 gpurun_out/bench_n$n.err
  python - $n <<'PY'
import json, sys
n = sys.argv[1]
d=json.loads(open('gpurun_out/bench_n%s.json' % n).read().strip().splitlines()[-1])
print('N=%s value %.0f ms/step %.1f scan_kernel %.1f e2e %.0f (%.1f ms) S=%s eigh %s'%(n, d['value'], d['ms_per_step'], d['roofline']['launch_ms'], d['e2e']['value'], d['e2e']['ms_per_step'], d['roofline']['slices'], d['eigh_seconds']))
print(' stages', {k: round(1e3*v,2) for k,v in d['stage_seconds_per_step'].items() if v})
print(' e2e stages', {k: round(1e3*v,2) for k,v in d['e2e']['stage_seconds_per_step'].items() if v}, d['e2e']['h2d_lanes'])
PY
  for r in $(seq 0 $((n-1))); do cat gpurun_out/bench_rank$r.json; echo; done
done
