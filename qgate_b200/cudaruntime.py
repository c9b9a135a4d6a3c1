"""`qgate.simulator.cudaruntime` replacement: the runtime module backed by the sm_100a engine.

Same module-level protocol as qgate/simulator/cudaruntime.py:24-91
(create_qubit_states, create_qubits_states_getter, set_preference, reset_preference,
module_init, module_finalize, initialized) so `Simulator(cudaruntime, **prefs)` works
with either this package's front end or the reference's.

There is NO fallback: if the CUDA library is missing or no device is visible the
import / first allocation raises.
"""
import atexit
import os
import sys

from . import _capi
from .native import RuntimeModule

LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'lib', 'libqgate_b200.so')


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError('libqgate_b200.so is not built ({}); run `python -c "import '
                          '__graft_entry__ as g; g.build()"` or qgate_b200/csrc/build.py.'
                          .format(LIB_PATH))
    api = _capi.CApi(LIB_PATH)
    if api.backend_name != 'cuda-sm_100a':
        raise ImportError('unexpected backend {} in {}'.format(api.backend_name, LIB_PATH))
    return api


_module = RuntimeModule(_load)
this = sys.modules[__name__]


def get_api():
    return _module.api


def set_preference(device_ids=[], max_po2idx_per_chunk=-1, memory_store_size=-1):
    _module.set_preference(device_ids, max_po2idx_per_chunk, memory_store_size)


def reset_preference(device_ids=[], max_po2idx_per_chunk=-1, memory_store_size=-1):
    _module.reset_preference()


def module_init():
    _module.module_init()


def module_finalize():
    _module.module_finalize()


def create_qubit_states(dtype):
    return _module.create_qubit_states(dtype)


def create_qubits_states_getter(dtype):
    return _module.create_qubits_states_getter(dtype)


def __getattr__(name):
    if name == 'api':
        return _module.api
    if name in ('initialized', 'device_ids', 'max_po2idx_per_chunk', 'memory_store_size',
                'native_instances'):
        return getattr(_module, name)
    if name in ('native_multi_qubit_ops', 'native_sample'):
        return True
    raise AttributeError(name)


atexit.register(module_finalize)
