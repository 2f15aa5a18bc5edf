import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a B200 (run on the GPU box with -m gpu)')


def golden(name):
    return np.load(os.path.join(GOLDEN, name))


@pytest.fixture(scope='session')
def built():
    """The in-tree shared library and the host check program (compiles them if needed; nvcc cross-compiles)."""
    import __graft_entry__ as g
    g.build()
    return g


@pytest.fixture(scope='session')
def ctx(built):
    from mixmogam_b200 import get_context
    return get_context(0)


def neglog10_rel_err(p, p_ref, floor=1e-3):
    """|d(-log10 p)| / max(-log10 p_ref, floor): the parity measure of BASELINE.json."""
    a, b = -np.log10(np.asarray(p, dtype=np.float64)), -np.log10(np.asarray(p_ref, dtype=np.float64))
    ok = np.isfinite(a) & np.isfinite(b)
    assert np.array_equal(np.isfinite(a), np.isfinite(b))
    return np.max(np.abs(a[ok] - b[ok]) / np.maximum(b[ok], floor)) if ok.any() else 0.0
