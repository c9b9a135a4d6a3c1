"""TEST INFRASTRUCTURE ONLY — runtime module over the UNMODIFIED reference CPU runtime.

oracle/_ref/libqgate_ref_cpu.so is the reference's own CPUQubitProcessor /
CPUQubitsStatesGetter / CPUSamplingPool (qgate/simulator/src/CPU*.cpp) compiled in place
from /root/reference by oracle/Makefile behind the C ABI of include/qgate_b200.h.  This
module exposes it through the same runtime-module protocol as qgate_b200.cudaruntime, so
a test can run one circuit on both and compare:

    sim_ref = qgate_b200.simulator.with_runtime(oracle.ref_runtime.module, dtype=np.float64)

Parity status: this IS the reference (kind "reference" in bench.py's cpu_baseline).
"""
import os
import subprocess

from qgate_b200 import _capi
from qgate_b200.native import RuntimeModule

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, '_ref', 'libqgate_ref_cpu.so')
REFERENCE_ROOT = '/root/reference'


def build(force=False):
    """Compile oracle/_ref from the reference sources when they are present (this
    container); on the GPU box the prebuilt .so that travelled with the snapshot is used."""
    if os.path.isdir(os.path.join(REFERENCE_ROOT, 'qgate', 'simulator', 'src')):
        if force:
            subprocess.check_call(['make', '-C', HERE, 'clean'], stdout=subprocess.DEVNULL)
        subprocess.check_call(['make', '-C', HERE, '-j8'], stdout=subprocess.DEVNULL,
                              stderr=subprocess.DEVNULL)
    return os.path.exists(LIB_PATH)


def available():
    return os.path.exists(LIB_PATH)


def _load():
    api = _capi.CApi(LIB_PATH)
    assert api.backend_name == 'reference-cpu'
    return api


module = RuntimeModule(_load)


def simulator(**prefs):
    from qgate_b200.simulator import with_runtime
    return with_runtime(module, **prefs)
