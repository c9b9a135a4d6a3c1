"""`qgate_b200.simulator.cuda(**prefs)` — drop-in for `qgate.simulator.cuda(**prefs)`
(qgate/simulator/__init__.py:26-28).  prefs: dtype=np.float32|np.float64,
circuit_prep=prefs.dynamic|static|one_static."""
from . import qubits
from .qubits import null, abs2, prob
from .simulator import Simulator


def cuda(**prefs):
    """One process drives one GPU.  Inside an initialised torch.distributed job (torchrun, one
    rank per GPU) big state vectors are sharded over the ranks (qgate_b200/dist.py)."""
    import sys
    from .. import cudaruntime
    td = sys.modules.get('torch.distributed')
    if td is not None and td.is_available() and td.is_initialized() and td.get_world_size() > 1:
        from .. import dist
        return Simulator(dist.runtime(cudaruntime, **prefs.pop('sharding', {})), **prefs)
    return Simulator(cudaruntime, **prefs)


def with_runtime(runtime_module, **prefs):
    """Simulator over any module/object implementing the runtime protocol (tests use this
    with the reference-CPU oracle runtime)."""
    return Simulator(runtime_module, **prefs)
