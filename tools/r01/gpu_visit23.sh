#!/bin/bash
# split-K Gram tail, pinned result buffers, cluster-8 multicast for the scan, full-size property test
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo build failed; tail -20 gpurun_out/build.log; exit 1; }
timeout 600 python -m pytest tests/test_gpu_kinship.py tests/test_gpu_reml_scan.py tests/test_gpu_full_size.py -x -q -m gpu -k "split_k or cluster_sizes or full_size or streamed_from_host" -p no:cacheprovider --timeout 400 --durations=8 > gpurun_out/tests_new.log 2>&1
echo "new tests rc=$?"; tail -16 gpurun_out/tests_new.log
for cs in 2 4 8; do
  MMG_SCAN_CLUSTER=$cs timeout 300 python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-e2e > gpurun_out/bench_cs$cs.json 2> gpurun_out/bench_cs$cs.err
  echo "cs=$cs rc=$?"; python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_cs$cs.json'))
    print('cs=$cs value %.0f ms/step %.1f scan_kernel %.2f gram %.2f frac_issue %.3f stages %s'%(d['value'], d['ms_per_step'], d['roofline']['launch_ms'], d['kinship']['gram_ms'], d['roofline']['frac_of_issue_rate'], {k: round(1e3*v,1) for k,v in d['stage_seconds_per_step'].items() if v}))
except Exception as e:
    print('parse failed', e)
PY
  tail -2 gpurun_out/bench_cs$cs.err
done
MMG_GRAM_SPLITK=0 timeout 300 python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-e2e > gpurun_out/bench_nosplit.json 2> gpurun_out/bench_nosplit.err
python -c "
import json
d=json.load(open('gpurun_out/bench_nosplit.json')); print('no split-K: gram %.2f ms'%d['kinship']['gram_ms'])"
