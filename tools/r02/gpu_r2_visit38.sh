#!/bin/bash
# The reference arm as the driver launches it for N > 1 (torchrun, rank 0 alone works, the others exit 0).
mkdir -p gpurun_out
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 > gpurun_out/bench_ref_n2.json 2> gpurun_out/bench_ref_n2.err; echo "ref N=2 rc=$?"
python -c "
import json
l=[x for x in open('gpurun_out/bench_ref_n2.json').read().strip().splitlines() if x.startswith('{')]
d=json.loads(l[-1]); print(len(l), 'json line(s):', d['impl'], d['value'], d['cpu_baseline']['cores'], d['n_gpus'], d['config']['workload'][:40])"
