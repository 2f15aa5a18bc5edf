#!/bin/bash
# N-GPU bench under torchrun, as the driver launches it
N=${N:-2}; M=${M:-1000000}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
   bench.py --gpus $N --steps 2 --warmup 1 --snps $M > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "bench N=$N rc=$?"; cat gpurun_out/bench_n$N.json; tail -5 gpurun_out/bench_n$N.err
