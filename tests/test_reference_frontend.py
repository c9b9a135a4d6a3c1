"""CPU suite, part 4 (build container only): the drop-in claim at the Python boundary.

The REAL reference package is imported from a scratch copy of /root/reference (recipe of
tests/golden/make_golden.py), this package's runtime objects (qgate_b200/native.py) are installed
as `qgate.simulator.cudaruntime` (qgate_b200/install.py), and the reference's OWN unittest
modules are run: their `...CUDA` classes then execute the reference's unmodified front end
(Simulator, ModelExecutor, RopExecutor, QubitsHandler, Qubits) on top of our wrappers and the C ABI.
Here the library behind the ABI is the reference-CPU shim (no GPU in this container); on a GPU box
the same wrappers sit on libqgate_b200.so.  Skipped where /root/reference is absent."""
import os
import sys
import types
import unittest

import numpy as np
import pytest

REFERENCE = '/root/reference'
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REFERENCE, 'tests')),
                                reason='/root/reference is not present (GPU box)')

MODULES = ['test_unary_gate', 'test_control_gate', 'test_calc_prob', 'test_measure', 'test_reset',
           'test_if', 'test_get_states', 'test_join', 'test_sampling_pool', 'test_swap_gate',
           'test_exp_gate', 'test_qreg_ordering', 'test_simple_calls',
           'test_z_conv', 'test_big_circuits', 'test_multidevice']
# (test_sampling, test_U_gate, test_gate_matrix, test_repr, test_pauli_*, test_memstore generate no
#  per-runtime ...CUDA classes: they exercise the reference's front end / py runtime only)


@pytest.fixture(scope='module')
def reference_with_our_runtime():
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden'))
    import make_golden
    qgate = make_golden.import_reference()
    from oracle import ref_runtime
    if not ref_runtime.available():
        ref_runtime.build()
    import qgate_b200.install
    qgate_b200.install.install(qgate, ref_runtime.module)
    if not hasattr(np, 'int'):
        np.int = int                  # the reference tests use the removed alias
    pkg = types.ModuleType('tests')   # skip tests/__init__.py (it star-imports the PLY parser tests)
    pkg.__path__ = [os.path.join(make_golden.SCRATCH, 'tests')]
    saved = sys.modules.get('tests')
    sys.modules['tests'] = pkg
    yield qgate
    if saved is not None:
        sys.modules['tests'] = saved
    for name in [n for n in sys.modules if n.startswith('tests.test_')]:
        if getattr(sys.modules[name], '__file__', '').startswith(make_golden.SCRATCH):
            del sys.modules[name]


@pytest.mark.parametrize('module', MODULES)
def test_reference_suite_on_our_runtime_objects(reference_with_our_runtime, module):
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
