#!/bin/bash
# One GPU-box visit: diagnostics, the gpu test-suite in isolated processes (a wedged tcgen05 kernel must
# not take the other results with it), a short bench and the ncu launch list.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi > gpurun_out/nvsmi.txt 2>&1
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 300 python tools/diag_tc.py > gpurun_out/diag_tc.log 2>&1; echo "diag_tc rc=$?"
run() { name=$1; shift; timeout 900 python -m pytest "$@" -q -m gpu -p no:cacheprovider --timeout 600 > gpurun_out/$name.log 2>&1; echo "$name rc=$?"; tail -3 gpurun_out/$name.log; }
run kin_simt tests/test_gpu_kinship.py -k "not tcgen05"
run kin_tc tests/test_gpu_kinship.py -k "tcgen05"
MMG_SCAN_IMPL=dmma run scan_dmma tests/test_gpu_reml_scan.py -k "not tcgen05 and not agree"
run scan_tc tests/test_gpu_reml_scan.py -k "tcgen05 or agree"
MMG_SCAN_IMPL=dmma run hdf5 tests/test_gpu_hdf5.py
