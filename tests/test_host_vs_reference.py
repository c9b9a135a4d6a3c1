"""CPU suite, part 1: the Python host stack (script API, preprocessor / dynamic grouping,
simulator, Qubits slicing, sampling-pool plumbing, ctypes binding) and the C-ABI shim over
the UNMODIFIED reference CPU runtime reproduce, bit for bit, what the real reference
(qgate.simulator.cpu driven by qgate's own front end) recorded in tests/golden/.

This pins the oracle the GPU tests compare against: same native code, so any difference
would be a host-side (lane bookkeeping, RNG draw order, argument marshalling) bug."""
import numpy as np
import pytest

import qgate_b200.script as S
from tests import cases

PREPS = ('dynamic', 'one_static')
RTS = (('cpu64', np.float64), ('cpu32', np.float32))
COLLAPSE_TOL = {np.float64: 1e-14, np.float32: 1e-6}


@pytest.mark.parametrize('name', sorted(cases.CIRCUITS))
@pytest.mark.parametrize('prep', PREPS)
@pytest.mark.parametrize('rt,dtype', RTS)
def test_states_prob_p0_bit_exact(golden, ref_runtime, name, prep, rt, dtype):
    sim, q = cases.run_circuit(ref_runtime.module, name, dtype, prep)
    key = '{}/{}/{}'.format(name, prep, rt)
    states = sim.qubits.states[:]
    assert states.dtype == golden[key + '/states'].dtype
    assert np.array_equal(states, golden[key + '/states'])
    assert np.array_equal(sim.qubits.prob[:], golden[key + '/prob'])
    p0 = np.array([sim.qubits.calc_probability(qr) for qr in q])
    # the reference's calc_probability depends on its worker count at ~1e-14
    assert np.allclose(p0, golden[key + '/p0'], rtol=0, atol=1e-12 if dtype is np.float64 else 1e-5)


@pytest.mark.parametrize('name', sorted(cases.CIRCUITS))
def test_cpu64_matches_py_runtime(golden, name):
    """The two reference runtimes agree with each other in the golden file (sanity of the
    fixtures themselves): py is always complex128."""
    for prep in PREPS:
        a = golden['{}/{}/py/states'.format(name, prep)]
        b = golden['{}/{}/cpu64/states'.format(name, prep)]
        assert cases.rel_err(b, a) < 1e-13


@pytest.mark.parametrize('name', ('rand6x10', 'zoo7', 'grover8', 'ghz11'))
@pytest.mark.parametrize('prep', PREPS)
@pytest.mark.parametrize('rt,dtype', RTS)
def test_measure_all_bits_and_collapsed_state(golden, ref_runtime, name, prep, rt, dtype):
    sim = cases.make_sim(ref_runtime.module, dtype, prep)
    q, ops = cases.CIRCUITS[name]()
    refs = S.new_references(len(q))
    np.random.seed(11)
    sim.run(ops + [S.measure(r, qr) for r, qr in zip(refs, q)])
    sim.qubits.set_ordering(q)
    key = 'measure/{}/{}/{}'.format(name, prep, rt)
    assert np.array_equal(np.array(sim.values.get(refs), np.int64), golden[key + '/bits'])
    # P(0) is a reduction whose order follows the lane layout, and the layout follows python
    # set iteration over qreg ids (qubits_handler.py:45-56) -> the 1/sqrt(p) factor may differ
    # in the last bit between two runs of the reference itself.
    assert cases.rel_err(sim.qubits.states[:], golden[key + '/states']) < COLLAPSE_TOL[dtype]
    assert int(sim.obs(refs)) == sum(int(b) << i for i, b in enumerate(golden[key + '/bits']))


@pytest.mark.parametrize('prep', PREPS)
@pytest.mark.parametrize('rt,dtype', RTS)
def test_midcircuit_measure_reset_if(golden, ref_runtime, prep, rt, dtype):
    sim = cases.make_sim(ref_runtime.module, dtype, prep)
    q, refs, ops = cases.midcircuit_ops()
    np.random.seed(5)
    sim.run(ops)
    sim.qubits.set_ordering(q)
    key = 'measure/midcircuit/{}/{}'.format(prep, rt)
    assert np.array_equal(np.array(sim.values.get(refs), np.int64), golden[key + '/bits'])
    assert cases.rel_err(sim.qubits.states[:], golden[key + '/states']) < COLLAPSE_TOL[dtype]


@pytest.mark.parametrize('name', ('rand10x20', 'zoo9', 'grover8'))
@pytest.mark.parametrize('prep', PREPS)
@pytest.mark.parametrize('rt,dtype', RTS)
def test_sampling_pool_indices(golden, ref_runtime, name, prep, rt, dtype):
    rnd = np.random.RandomState(7).random_sample(20000)
    sim = cases.make_sim(ref_runtime.module, dtype, prep)
    q, ops = cases.CIRCUITS[name]()
    empty = S.new_qregs(2)
    sim.run(ops)
    key = 'sampling/{}/{}/{}'.format(name, prep, rt)
    orderings = {'full': q, 'hidden': q[1::2],
                 'empty': [q[3], empty[0], q[0], q[5], empty[1], q[1]],
                 'reversed': list(reversed(q))}
    for tag, ordering in orderings.items():
        got = sim.qubits.create_sampling_pool(ordering).sample(20000, rnd).intarray
        assert np.array_equal(got, golden[key + '/' + tag]), tag

    class Probe:
        def __init__(self, prob, empty_lanes, qreg_ordering):
            self.prob = prob
    prob = sim.qubits.create_sampling_pool(q[1::2], Probe).prob
    # hidden lanes are summed in local-lane order, which follows the (id-dependent) layout
    assert np.allclose(prob, golden[key + '/prob_hidden'], rtol=1e-13 if dtype is np.float64 else 1e-5, atol=0)


@pytest.mark.parametrize('rt,dtype', RTS)
def test_get_states_slices(golden, ref_runtime, rt, dtype):
    sim, q = cases.run_circuit(ref_runtime.module, 'rand10x20', dtype, 'dynamic')
    extra = S.new_qregs(1)
    for idx, k in enumerate(cases.SLICE_KEYS):
        assert np.array_equal(sim.qubits.states[slice(*k)],
                              golden['slices/{}/{}/states'.format(rt, idx)]), k
        assert np.array_equal(sim.qubits.prob[slice(*k)],
                              golden['slices/{}/{}/prob'.format(rt, idx)]), k
    sim.qubits.set_ordering(q[:4] + extra + q[4:])
    assert np.array_equal(sim.qubits.states[::5], golden['slices/{}/empty_lane/states'.format(rt)])
    # scalar indexing
    sim.qubits.set_ordering(q)
    full = golden['slices/{}/0/states'.format(rt)]
    assert sim.qubits.states[5] == full[5]
    assert sim.qubits.states[-1] == full[-1]
    with pytest.raises(RuntimeError):
        sim.qubits.states[1 << 10]


@pytest.mark.parametrize('rt,dtype', RTS)
def test_phase_estimation_hidden_target(golden, ref_runtime, rt, dtype):
    from qgate_b200 import circuits
    sim = cases.make_sim(ref_runtime.module, dtype, 'dynamic')
    bits, target, ops = circuits.phase_estimation(S, 8, 0.1)
    sim.run(ops)
    sim.qubits.set_ordering(bits + [target])
    assert np.array_equal(sim.qubits.states[:], golden['pe8/{}/states'.format(rt)])
    rnd = np.random.RandomState(3).random_sample(4096)
    got = sim.qubits.create_sampling_pool(bits).sample(4096, rnd).intarray
    assert np.array_equal(got, golden['pe8/{}/samples'.format(rt)])


def test_gate_matrices_match_reference_columns(golden, ref_runtime):
    api = ref_runtime.module.api
    from qgate_b200 import _capi
    for name, params in cases.GATE_SPECS:
        gid = _capi.GATE_IDS[cases.SCRIPT_TO_GATE_ID.get(name, name)]
        for adj in (False, True):
            mat = api.gate_matrix(gid, params, adj)
            want = golden['matrix/{}/{}'.format(name, 'adj' if adj else 'fwd')]
            assert np.array_equal(mat, want), (name, adj)


def test_qft20_config0(golden, ref_runtime):
    """BASELINE.json configs[0]: 20-qubit QFT, complex128, against the reference's cpu run
    (and its py run when the fixture holds it)."""
    from qgate_b200 import circuits
    sim = cases.make_sim(ref_runtime.module, np.float64, 'one_static')
    q, ops = circuits.qft(S, 20)
    sim.run(ops)
    sim.qubits.set_ordering(q)
    assert np.array_equal(sim.qubits.states[::4099], golden['qft20/cpu64/states_stride4099'])
    assert np.array_equal(sim.qubits.states[:512], golden['qft20/cpu64/states_head'])
    if 'qft20/py/states_head' in golden:
        assert cases.rel_err(sim.qubits.states[:512], golden['qft20/py/states_head']) < 1e-12
        assert cases.rel_err(sim.qubits.states[::4099],
                             golden['qft20/py/states_stride4099']) < 1e-12


def test_more_than_40_separate_qregs_on_the_reference_runtime(ref_runtime):
    """the same product-state case the GPU suite runs (tests/test_gpu_parity.py), on the reference CPU
    runtime: pins the expectations of that test (Rz convention, index order, pool over a subset)"""
    from tests.test_gpu_parity import _product_state_case
    _product_state_case(ref_runtime.module, np.float64)
