"""`qgate_b200.simulator.cuda(**prefs)` — drop-in for `qgate.simulator.cuda(**prefs)`
(qgate/simulator/__init__.py:26-28).  prefs: dtype=np.float32|np.float64,
circuit_prep=prefs.dynamic|static|one_static."""
from . import qubits
from .qubits import null, abs2, prob
from .simulator import Simulator


def cuda(**prefs):
    from .. import cudaruntime
    return Simulator(cudaruntime, **prefs)


def with_runtime(runtime_module, **prefs):
    """Simulator over any module/object implementing the runtime protocol (tests use this
    with the reference-CPU oracle runtime)."""
    return Simulator(runtime_module, **prefs)
