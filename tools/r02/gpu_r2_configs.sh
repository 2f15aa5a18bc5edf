#!/bin/bash
# Round 2: the BASELINE.json configurations other than configs[1] on one B200 (tools/bench_configs.py), and the n = 10 000
# parity run against the line-faithful oracle (MMG_TEST_ORACLE_FULL=1, m cut to 131072 so the visit stays short).
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo build failed; tail -20 gpurun_out/build.log; exit 1; }
nvidia-smi --query-gpu=name,memory.total,memory.used --format=csv > gpurun_out/nvsmi.txt
for c in ${CONFIGS:-2 4 3}; do
  timeout ${CFG_TIMEOUT:-900} python tools/bench_configs.py --config $c > gpurun_out/r02_config$c.json 2> gpurun_out/r02_config$c.err
  echo "config $c rc=$?"; tail -c 1500 gpurun_out/r02_config$c.json; tail -3 gpurun_out/r02_config$c.err
done
if [ -n "$PARITY" ]; then
MMG_TEST_FULL_M=131072 MMG_TEST_ORACLE_FULL=1 timeout 1200 python -m pytest tests/test_gpu_full_size.py -q -s -p no:cacheprovider > gpurun_out/r02_parity_n10k.txt 2>&1
echo "parity rc=$?"; grep -a "oracle(double)\|passed\|failed" gpurun_out/r02_parity_n10k.txt
fi
