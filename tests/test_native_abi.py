"""CPU suite, part 2: the product library itself, without a GPU.

* libqgate_b200.so loads and exports every symbol include/qgate_b200.h declares,
* it reports the CUDA backend and refuses to compute without a device (no CPU fallback),
* its host-side gate-matrix factory reproduces the reference's 16 matrices (golden vectors
  recorded from the real reference, GateMatrix.cpp:13-168),
* the flush planner is exercised by a CPU emulator of the tile kernel (tests/native).
"""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

from qgate_b200 import _capi, cudaruntime
from tests import cases

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def lib_path():
    if not os.path.exists(cudaruntime.LIB_PATH):
        from qgate_b200.csrc import build
        build.build()
    return cudaruntime.LIB_PATH


def header_symbols():
    text = open(os.path.join(REPO, 'include', 'qgate_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(qgb_[a-z0-9_]+)\s*\(', text)))


def test_header_and_binding_agree():
    assert header_symbols() == _capi.exported_symbols()


def test_library_exports_every_header_symbol(lib_path):
    lib = ctypes.CDLL(lib_path)
    for name in header_symbols():
        assert hasattr(lib, name), name


def test_backend_is_cuda_and_there_is_no_cpu_fallback(lib_path):
    api = _capi.CApi(lib_path)
    assert api.backend_name == 'cuda-sm_100a'
    assert api.lib.qgb_abi_version() == 1
    if api.device_count() > 0:
        pytest.skip('a GPU is present; the no-device behaviour is checked on the CPU box')
    ids = api.int_array([])
    with pytest.raises(RuntimeError):
        api.call('qgb_devices_initialize', ids, 0, -1, -1)
    # objects can be created, but nothing computes without a device
    h = api.new_handle('qgb_qstates_new', _capi.PREC_FP64)
    p = api.new_handle('qgb_qproc_new', _capi.PREC_FP64)
    with pytest.raises(RuntimeError):
        api.call('qgb_qproc_initialize_qstates', p, h, 4)
    api.call('qgb_qstates_delete', h)
    api.call('qgb_qproc_delete', p)
    with pytest.raises(ValueError):
        api.call('qgb_qstates_delete', h)


def test_gate_matrices_match_reference(golden, lib_path):
    api = _capi.CApi(lib_path)
    for name, params in cases.GATE_SPECS:
        gid = _capi.GATE_IDS[cases.SCRIPT_TO_GATE_ID.get(name, name)]
        for adj in (False, True):
            mat = api.gate_matrix(gid, params, adj)
            want = golden['matrix/{}/{}'.format(name, 'adj' if adj else 'fwd')]
            assert np.array_equal(mat, want), (name, adj, mat, want)
    with pytest.raises(RuntimeError):
        api.gate_matrix(99, (), False)
    with pytest.raises(ValueError):
        api.gate_matrix(_capi.GATE_IDS['U'], (0.1,), False)


def test_planner_against_tile_kernel_emulator(tmp_path):
    exe = str(tmp_path / 'planner_emul')
    csrc = os.path.join(REPO, 'qgate_b200', 'csrc')
    subprocess.check_call(['g++', '-std=c++17', '-O2', '-I' + csrc,
                           os.path.join(REPO, 'tests', 'native', 'planner_emul.cpp'),
                           os.path.join(csrc, 'planner.cpp'), '-o', exe])
    out = subprocess.run([exe], stdout=subprocess.PIPE, text=True)
    assert out.returncode == 0, out.stdout[-4000:]
    assert 'ALL OK' in out.stdout


def test_memory_pool_against_a_mock_allocator(tmp_path):
    """engine.cu's MemPool (size classes, caps, eviction, trim-and-retry, the no-free contract for blocks
    exported through CUDA IPC) compiled as it stands in the product against a mock cudaMalloc / cudaFree
    with a fixed capacity: tests/native/pool_emul.cpp."""
    src = open(os.path.join(REPO, 'qgate_b200', 'csrc', 'engine.cu')).read()
    begin, end = src.index('struct MemPool {'), src.index('/* ---- objects')
    with open(str(tmp_path / 'pool_extract.h'), 'w') as f:
        f.write(src[begin:end])
    exe = str(tmp_path / 'pool_emul')
    subprocess.check_call(['g++', '-std=c++17', '-O1', '-I' + str(tmp_path),
                           os.path.join(REPO, 'tests', 'native', 'pool_emul.cpp'), '-o', exe])
    out = subprocess.run([exe], stdout=subprocess.PIPE, text=True)
    assert out.returncode == 0, out.stdout[-4000:]
    assert 'ALL OK' in out.stdout
