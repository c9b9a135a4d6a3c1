"""The drop-in claim at the Python boundary.

The REAL reference package is imported (baseline/reference_frontend.py: a scratch build of
/root/reference in the build container, baseline/_ref on the GPU box), this package's runtime objects (qgate_b200/native.py) are installed
as `qgate.simulator.cudaruntime` (qgate_b200/install.py), and the reference's OWN unittest
modules are run: their `...CUDA` classes then execute the reference's unmodified front end
(Simulator, ModelExecutor, RopExecutor, QubitsHandler, Qubits) on top of our wrappers and the C ABI.
The library behind the ABI is the reference-CPU shim in the CPU suite and libqgate_b200.so in the
`-m gpu` suite (test_reference_suite_on_the_cuda_library)."""
import os
import sys
import types
import unittest

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from baseline import reference_frontend  # noqa: E402

pytestmark = pytest.mark.skipif(not reference_frontend.available(),
                                reason='the reference package is neither under baseline/_ref nor at /root/reference')

MODULES = ['test_unary_gate', 'test_control_gate', 'test_calc_prob', 'test_measure', 'test_reset',
           'test_if', 'test_get_states', 'test_join', 'test_sampling_pool', 'test_swap_gate',
           'test_exp_gate', 'test_qreg_ordering', 'test_simple_calls',
           'test_z_conv', 'test_big_circuits', 'test_multidevice']
# (test_sampling, test_U_gate, test_gate_matrix, test_repr, test_pauli_*, test_memstore generate no
#  per-runtime ...CUDA classes: they exercise the reference's front end / py runtime only)


@pytest.fixture(scope='module')
def reference_package():
    qgate = reference_frontend.load()
    pkg = types.ModuleType('tests')   # skip tests/__init__.py (it star-imports the PLY parser tests)
    pkg.__path__ = [os.path.join(reference_frontend.root(), 'tests')]
    saved = sys.modules.get('tests')
    sys.modules['tests'] = pkg
    yield qgate
    if saved is not None:
        sys.modules['tests'] = saved
    for name in [n for n in sys.modules if n.startswith('tests.test_')]:
        if getattr(sys.modules[name], '__file__', '').startswith(reference_frontend.root()):
            del sys.modules[name]


@pytest.fixture
def reference_with_our_runtime(reference_package):
    """the reference front end over this package's wrappers + the reference-CPU shim (no GPU needed)"""
    from oracle import ref_runtime
    if not ref_runtime.available():
        ref_runtime.build()
    import qgate_b200.install
    qgate_b200.install.install(reference_package, ref_runtime.module)
    return reference_package


@pytest.fixture
def reference_on_cuda_library(reference_package, cuda_runtime):
    """the reference front end over this package's wrappers + libqgate_b200.so (needs a B200)"""
    import qgate_b200.install
    qgate_b200.install.install(reference_package, cuda_runtime)
    return reference_package


def run_reference_module(module):
    mod = __import__('tests.' + module, fromlist=['*'])
    loader = unittest.TestLoader()
    suite = unittest.TestSuite()
    n_classes = 0
    for name in dir(mod):
        obj = getattr(mod, name)
        if isinstance(obj, type) and issubclass(obj, unittest.TestCase) and name.endswith('CUDA'):
            suite.addTests(loader.loadTestsFromTestCase(obj))
            n_classes += 1
    assert n_classes > 0, 'no ...CUDA test classes were generated: cudaruntime was not picked up'
    result = unittest.TestResult()
    suite.run(result)
    problems = ['{}: {}'.format(t.id(), tb.splitlines()[-1]) for t, tb in result.errors + result.failures]
    assert not problems, '\n'.join(problems)
    assert result.testsRun > 0
    return result.testsRun


@pytest.mark.parametrize('module', MODULES)
def test_reference_suite_on_our_runtime_objects(reference_with_our_runtime, module):
    run_reference_module(module)


@pytest.mark.gpu
@pytest.mark.parametrize('module', MODULES)
def test_reference_suite_on_the_cuda_library(reference_on_cuda_library, module):
    """The drop-in claim on the real thing: the reference's unmodified front end and its own
    unittest classes, `qgate.simulator.cuda()` served by libqgate_b200.so on the B200."""
    run_reference_module(module)
    api = reference_on_cuda_library.simulator.cudaruntime.get_api()
    assert api.backend_name == 'cuda-sm_100a'


def test_reference_openqasm_tests_on_our_reader(reference_with_our_runtime):
    """the reference's own tests/test_openqasm.py (statement forms and error
    behaviour) against `qgate.openqasm` as installed by qgate_b200.install (the reference's importer
    needs PLY, which is not in this image), plus examples/qft.qasm run on the reference's cpu runtime
    against the same circuit written with the reference's script API."""
    qgate = reference_with_our_runtime
    assert getattr(qgate.openqasm, '_qgate_b200', False)
    mod = __import__('tests.test_openqasm', fromlist=['*'])
    suite = unittest.TestLoader().loadTestsFromTestCase(mod.TestOpenQASM)
    result = unittest.TestResult()
    suite.run(result)
    problems = ['{}: {}'.format(t.id(), tb.splitlines()[-1]) for t, tb in result.errors + result.failures]
    assert not problems, '\n'.join(problems)
    assert result.testsRun >= 19
    import importlib
    import math
    RS = importlib.import_module('qgate.script')
    prog = qgate.openqasm.load_circuit_from_file(os.path.join(reference_frontend.root(), 'examples', 'qft.qasm'))
    gates = [op for op in prog.circuit if not isinstance(op, qgate.model.Measure)]
    sim = qgate.simulator.cpu(dtype=np.float64)
    sim.run([RS.X(prog.q[1])] + gates)
    sim.qubits.set_ordering(prog.q)
    got = sim.qubits.states[:]
    q = RS.new_qregs(4)
    ops = [RS.X(q[1]), RS.H(q[3]), RS.ctrl(q[2]).U1(math.pi / 2)(q[3]), RS.ctrl(q[1]).U1(math.pi / 4)(q[3]),
           RS.ctrl(q[0]).U1(math.pi / 8)(q[3]), RS.H(q[2]), RS.ctrl(q[1]).U1(math.pi / 2)(q[2]),
           RS.ctrl(q[0]).U1(math.pi / 4)(q[2]), RS.H(q[1]), RS.ctrl(q[0]).U1(math.pi / 2)(q[1]), RS.H(q[0])]
    sim2 = qgate.simulator.cpu(dtype=np.float64)
    sim2.run(ops)
    sim2.qubits.set_ordering(q)
    assert np.abs(got - sim2.qubits.states[:]).max() < 1e-15


def test_simulator_sample_matches_reference_shot_for_shot(reference_with_our_runtime):
    """`Simulator.sample` (simulator.py:85-119: the whole circuit re-run per shot, one global-RNG
    draw per Measure): our front end on the oracle runtime and the real reference on its cpu
    runtime return the same observation for every shot under the same seed."""
    qgate = reference_with_our_runtime
    import qgate_b200
    import qgate_b200.script as S
    from oracle import ref_runtime

    def build(mod):
        q = mod.new_qregs(3)
        refs = mod.new_references(3)
        ops = [mod.H(q[0]), mod.ctrl(q[0]).X(q[1]), mod.Ry(0.7)(q[2]), mod.ctrl(q[1]).Rx(1.1)(q[2]),
               mod.measure(refs[0], q[0]), mod.measure(refs[1], q[1]),
               mod.if_(refs[1], 1, mod.X(q[2])), mod.measure(refs[2], q[2])]
        return refs, ops

    import importlib
    ref_script = importlib.import_module('qgate.script')
    refs, ops = build(ref_script)
    ref_sim = qgate.simulator.cpu(dtype=np.float64)
    np.random.seed(5)
    want = ref_sim.sample(ops, refs, 300).intarray
    refs, ops = build(S)
    sim = qgate_b200.simulator.with_runtime(ref_runtime.module, dtype=np.float64)
    np.random.seed(5)
    got = sim.sample(ops, refs, 300).intarray
    assert np.array_equal(got, want)
    assert len(set(got.tolist())) > 2
