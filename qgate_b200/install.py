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
import importlib
import sys
import types


def install(qgate_package, runtime_module=None):
    if runtime_module is None:
        from . import cudaruntime as runtime_module
    simulator = qgate_package.simulator
    simulator.cudaruntime = runtime_module
    sys.modules[simulator.__name__ + '.cudaruntime'] = runtime_module
    install_openqasm(qgate_package)
    return runtime_module


def install_openqasm(qgate_package):
    """`qgate.openqasm` (importer.py:9-50) needs PLY; where that import fails (or was stubbed out), serve
    its four entry points from this package's own OpenQASM reader, building with the REFERENCE's script
    API so the circuit objects are the reference's own."""
    name = qgate_package.__name__ + '.openqasm'
    existing = sys.modules.get(name)
    if existing is None:
        try:
            existing = importlib.import_module(name)
        except Exception:
            existing = None
    if existing is not None and hasattr(existing, 'load_circuit') and not getattr(existing, '_qgate_b200', False):
        return existing                       # the reference's own importer works here: keep it
    from . import openqasm as reader
    script = importlib.import_module(qgate_package.__name__ + '.script')
    mod = types.ModuleType(name)
    mod._qgate_b200 = True
    mod.translate = lambda qasm: reader.translate(qasm, qgate_package.__name__)
    mod.translate_file = lambda filename: reader.translate_file(filename, qgate_package.__name__)
    mod.load_circuit = lambda qasm: reader.load_circuit(qasm, script)
    mod.load_circuit_from_file = lambda filename: reader.load_circuit_from_file(filename, script)
    sys.modules[name] = mod
    qgate_package.openqasm = mod
    return mod
