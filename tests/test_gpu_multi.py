"""GPU (needs >= 2 GPUs, skipped otherwise): the SNP-sharded path end to end under torchrun with NCCL -- bit-exact kinship from
per-rank partial Grams, the quadratic form formed block-wise across the ranks, gathered per-SNP results, against the single-GPU
path and the FP64 oracle (tests/multi_gpu_worker.py does the work)."""
import json
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize('world', [2])
def test_sharded_emmax_matches_single_gpu_and_oracle(built, world):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip('needs %d GPUs' % world)
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', str(world), '--master-addr', '127.0.0.1',
           '--master-port', str(_free_port()), os.path.join(ROOT, 'tests', 'multi_gpu_worker.py')]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    line = [l for l in p.stdout.splitlines() if l.startswith('MULTI_RESULT ')][-1]
    out = json.loads(line[len('MULTI_RESULT '):])
    assert out['err_vs_oracle'] < 1e-6 and out['err_vs_single'] < 1e-7
    assert abs(out['her'][0] - out['her'][1]) < 1e-8
