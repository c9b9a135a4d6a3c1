"""Install this package's runtime as `qgate.simulator.cudaruntime` of an imported shinmorino/qgate.

    import qgate                                  # the reference package (0.2.x)
    import qgate_b200.install
    qgate_b200.install.install(qgate)             # qgate.simulator.cuda() now runs on the B200 engine

The reference's `qgate.simulator.cuda(**prefs)` is `Simulator(cudaruntime, **prefs)` with
`cudaruntime` looked up in `qgate.simulator`'s namespace (qgate/simulator/__init__.py:9-12,26-28), and
its test factory creates the `...CUDA` test classes when that attribute exists
(tests/test_base.py:47-53) — so assigning the attribute is the whole installation.  The runtime
objects (qgate_b200/native.py) accept the reference's own gate-type, lane and math-op objects.
"""
import sys


def install(qgate_package, runtime_module=None):
    if runtime_module is None:
        from . import cudaruntime as runtime_module
    simulator = qgate_package.simulator
    simulator.cudaruntime = runtime_module
    sys.modules[simulator.__name__ + '.cudaruntime'] = runtime_module
    return runtime_module
