import os
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# Pin the worker count of the reference CPU runtime (oracle/_ref reads QGATE_NUM_WORKERS once, at its
# first parallel loop: Parallel.cpp:21-38).  Its sampling pool scans one span per worker, so the
# count is part of what "the reference's sampled indices" means; tests that ask the engine for the
# reference-compatible pool (option pool_compat_workers) pass the same number.
os.environ.setdefault('QGATE_NUM_WORKERS', str(min(16, len(os.sched_getaffinity(0)))))
if REPO not in sys.path:
    sys.path.insert(0, REPO)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')


@pytest.fixture(scope='session')
def golden():
    import numpy as np
    path = os.path.join(REPO, 'tests', 'golden', 'golden_v1.npz')
    with np.load(path) as data:
        return {k: data[k] for k in data.files}


@pytest.fixture(scope='session')
def ref_runtime():
    """The unmodified reference CPU runtime behind the C ABI (oracle/_ref)."""
    from oracle import ref_runtime as rr
    if not rr.available():
        rr.build()
    if not rr.available():
        pytest.skip('oracle/_ref is not built and /root/reference is absent')
    return rr


@pytest.fixture(scope='session')
def cuda_runtime():
    import ctypes
    try:
        ctypes.CDLL('libcuda.so.1')
    except OSError:
        pytest.skip('no CUDA driver')
    from qgate_b200 import cudaruntime
    if cudaruntime.get_api().device_count() == 0:
        pytest.skip('no CUDA device')
    return cudaruntime


@pytest.fixture(autouse=True)
def _fresh_handle_ids():
    """Lane layout in dynamic mode follows python set iteration over qreg ids
    (qubits_handler.py:45-56), and float32 hidden-lane sums follow the layout: start every
    test from id 0 so results do not depend on which tests ran before."""
    import itertools
    from qgate_b200 import model
    model.Qreg._counter = itertools.count()
    model.Reference._counter = itertools.count()
    yield
