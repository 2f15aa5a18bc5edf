#!/bin/bash
# Round 2, visit 9: shared-rotation batch under different L2 eviction hints / K splits / clusters.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo build failed; tail -20 gpurun_out/build.log; exit 1; }
timeout 900 python -m pytest tests/test_gpu_reml_scan.py -q -m gpu -p no:cacheprovider -x -k "multi" > gpurun_out/t_new.log 2>&1; echo "t_new rc=$?"; tail -4 gpurun_out/t_new.log
timeout 900 python tools/bench_multi.py --indivs 10000 --snps 131072 --phenotypes 199 --single 1 --unshared 0 --envs "$ENVS" > gpurun_out/r02_multi.json 2> gpurun_out/r02_multi.err; echo "multi rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_multi.json').read().strip().splitlines()[-1])
print('default scan stage %.1f ms' % (1e3*d['stage_seconds']['scan']), d['shared_scan_info'])
for k,v in d.items():
    if k.startswith('env['): print(k, '%.1f ms' % (1e3*v['scan_stage_s']), 'rot %.1f con %.1f' % (v['info']['rotation_ms'], v['info']['contraction_ms']))
PY
tail -5 gpurun_out/r02_multi.err
