"""SURVEY.md section 8f rows f-1 (batch submit) and f-3 (native multi-qubit ops): Swap as a lane
relabelling, exp(i theta P) as basis changes around ONE parity-phase diagonal op, a whole gate list
in one native call — each against the reference's own expansion (qgate/model/expand.py:14-90)
executed gate by gate on the reference CPU runtime (oracle/_ref)."""
import numpy as np
import pytest

import qgate_b200
import qgate_b200.script as S
from qgate_b200 import circuits
from qgate_b200.native import NativeQubitProcessor
from qgate_b200.simulator import with_runtime


def multi_qubit_circuit(n, n_ops, seed):
    """dense gates, Swaps, (controlled) Pauli-string exponentials incl. repeated qregs and identities"""
    rng = np.random.RandomState(seed)
    q = S.new_qregs(n)
    ops = [S.H(x) for x in q] + [S.Ry(0.3 * (i + 1))(x) for i, x in enumerate(q)]
    paulis = (S.X, S.Y, S.Z, S.I)
    n_swap = n_exp = 0
    while len(ops) < n_ops:
        kind = rng.randint(4)
        lanes = [int(l) for l in rng.permutation(n)]
        if kind == 0:
            ops.append(S.Swap(q[lanes[0]], q[lanes[1]]))
            n_swap += 1
        elif kind == 1:
            m = int(rng.randint(1, min(6, n - 1)))
            string = [paulis[rng.randint(4)](q[l]) for l in lanes[:m]]
            if rng.randint(3) == 0:      # the same qreg twice with the same letter: still a real coefficient
                string.append(string[0].copy())
            theta = float(rng.uniform(-2., 2.))
            if rng.randint(2):
                ops.append(S.ctrl(q[lanes[m]]).Expi(theta)(string))
            else:
                ops.append(S.Expi(theta)(string))
            n_exp += 1
        elif kind == 2:
            ops.append(S.ctrl(q[lanes[0]]).U3(*[float(v) for v in rng.uniform(0., 6., 3)])(q[lanes[1]]))
        else:
            ops.append(S.U3(*[float(v) for v in rng.uniform(0., 6., 3)])(q[lanes[0]]))
    assert n_swap > 3 and n_exp > 3
    return q, ops


def run(runtime, dtype, prep, q, ops, native, tail=None):
    sim = with_runtime(runtime, dtype=dtype, circuit_prep=prep, native_multi_qubit_ops=native)
    np.random.seed(77)
    sim.run(ops + (tail or []))
    sim.qubits.set_ordering(q)
    out = sim.qubits.states[:], np.array([sim.qubits.calc_probability(x) for x in q])
    sim.terminate()
    return out


@pytest.mark.parametrize('prep', ('dynamic', 'one_static'))
def test_front_end_keeps_swap_and_expi_whole(ref_runtime, prep):
    """CPU: the front end hands Swap / Expi to the runtime unexpanded when it can take them, and the
    result equals the expanded circuit (here the runtime is the reference-CPU shim, whose entry points
    ARE the expansion: what is checked is the grouping, the lane bookkeeping, the Pauli reduction)."""
    q, ops = multi_qubit_circuit(7, 60, seed=4)
    refs = S.new_references(2)
    tail = [S.measure(refs[0], q[2]), S.Swap(q[2], q[5]), S.measure(refs[1], q[5])]
    a, pa = run(ref_runtime.module, np.float64, prep, q, ops, True, tail)
    b, pb = run(ref_runtime.module, np.float64, prep, q, ops, False, tail)
    assert np.abs(a - b).max() < 1e-13 and np.abs(pa - pb).max() < 1e-13


def test_batch_submit_equals_gate_by_gate_on_the_shim(ref_runtime):
    _batch_vs_single(ref_runtime.module, np.float64, 9, 1e-15)


def _batch_vs_single(runtime, dtype, n, tol):
    q, ops = circuits.mixed_gate_zoo(S, n, 300, seed=23)
    outs = []
    for batch in (False, True):
        qstates = runtime.create_qubit_states(dtype)
        proc = qstates.processor
        proc.initialize_qubit_states(qstates, n)
        proc.reset_qubit_states(qstates)
        gates = [(op.gate_type, op.adjoint, [c.id - q[0].id for c in (op.ctrllist or [])], op.qreg.id - q[0].id)
                 for op in ops]
        if batch:
            proc.apply_gates_batch(qstates, NativeQubitProcessor.pack_gates(gates))
        else:
            for gate_type, adjoint, ctrls, target in gates:
                if ctrls:
                    proc.apply_controlled_gate(gate_type, adjoint, qstates, ctrls, target)
                else:
                    proc.apply_gate(gate_type, adjoint, qstates, target)
        getter = runtime.create_qubits_states_getter(dtype)
        values = np.empty(1 << n, np.complex128 if dtype is np.float64 else np.complex64)

        class Lane:
            def __init__(self, l):
                self.local = self.external = l
        from qgate_b200.simulator import qubits as qb
        getter.get_states(values, 0, qb.null, [(qstates, [Lane(l) for l in range(n)])], [], 1 << n, 0, 1)
        outs.append(values)
        getter.delete()
        qstates.delete()
    assert np.abs(outs[0] - outs[1]).max() <= tol


@pytest.mark.gpu
@pytest.mark.parametrize('dtype', (np.float64, np.float32))
@pytest.mark.parametrize('prep', ('dynamic', 'one_static'))
@pytest.mark.parametrize('n,n_ops', ((6, 60), (12, 150), (17, 200)))
def test_native_swap_and_pauli_expi_match_reference_expansion(cuda_runtime, ref_runtime, dtype, prep, n, n_ops):
    """GPU: the engine's native Swap / exp(i theta P) against the reference's expansion on the reference
    CPU runtime: small states (one kernel per gate), fused passes (parity diagonals over register,
    thread and outside-tile lanes), measurements and further Swaps on relabelled lanes."""
    api = cuda_runtime.get_api()
    q, ops = multi_qubit_circuit(n, n_ops, seed=n)
    refs = S.new_references(2)
    tail = [S.measure(refs[0], q[2]), S.Swap(q[2], q[n - 1]), S.ctrl(q[1]).Expi(0.4)([S.Z(q[2]), S.Y(q[0])]),
            S.measure(refs[1], q[n - 1])]
    api.stats_reset()
    got, p_got = run(cuda_runtime, dtype, prep, q, ops, True, tail)
    stats = api.stats()
    assert stats['native_swaps'] > 3 and stats['native_pauli_exps'] > 3, stats
    want, p_want = run(ref_runtime.module, dtype, prep, q, ops, False, tail)
    exact, _ = run(ref_runtime.module, np.float64, prep, q, ops, False, tail)
    tol = 1e-12 if dtype is np.float64 else 1e-5
    scale = np.abs(exact).max()
    assert np.abs(got - exact).max() / scale < tol
    assert np.abs(got - want).max() / scale < tol
    assert np.abs(p_got - p_want).max() < tol
    # and the engine expanding them itself gives the same state
    again, _ = run(cuda_runtime, dtype, prep, q, ops, False, tail)
    assert np.abs(got - again).max() / scale < tol


@pytest.mark.gpu
@pytest.mark.parametrize('dtype', (np.float64, np.float32))
def test_batch_submit_equals_gate_by_gate(cuda_runtime, dtype):
    _batch_vs_single(cuda_runtime, dtype, 13, 0.)


@pytest.mark.gpu
def test_sampling_pool_and_readout_follow_relabelled_lanes(cuda_runtime, ref_runtime):
    """after native Swaps the lane map is not the identity: pools (incl. the scan that reads the
    amplitudes directly), marginals, slices and joins must go through it"""
    n = 14
    q, ops = circuits.random_u3_cx(S, n, 4, seed=6)
    ops += [S.Swap(q[0], q[9]), S.Swap(q[3], q[13]), S.Swap(q[9], q[4])]
    ops += circuits.random_u3_cx(S, n, 2, seed=7, qregs=q)[1]
    extra = S.new_qregs(2)
    ops += [S.H(extra[0]), S.ctrl(extra[0]).X(q[5]), S.Swap(extra[1], q[2])]
    order = q + extra
    rnd = np.random.RandomState(2).random_sample(50000)
    outs = []
    for rt, native in ((cuda_runtime, True), (ref_runtime.module, False)):
        sim = with_runtime(rt, dtype=np.float64, circuit_prep='dynamic', native_multi_qubit_ops=native)
        sim.run(ops)
        sim.qubits.set_ordering(order)
        outs.append((sim.qubits.states[::3], sim.qubits.prob[5:900:7],
                     sim.qubits.create_sampling_pool(order).sample(len(rnd), rnd).intarray,
                     sim.qubits.create_sampling_pool(order[1::2]).sample(len(rnd), rnd).intarray))
        sim.terminate()
    assert np.abs(outs[0][0] - outs[1][0]).max() < 1e-12
    assert np.abs(outs[0][1] - outs[1][1]).max() < 1e-12
    assert np.array_equal(outs[0][2], outs[1][2])
    assert np.array_equal(outs[0][3], outs[1][3])


# ---- f-2: Simulator.sample without the per-shot re-run ------------------------------------------

def _terminal_circuit(n=10):
    q, ops = circuits.random_u3_cx(S, n, 5, seed=31)
    fresh = S.new_qreg()                       # never touched by a gate: measured as |0>, still one draw
    refs = S.new_references(7)
    order = [q[4], q[0], fresh, q[9], q[3], q[7], q[5]]
    ops = ops + [S.measure(r, x) for r, x in zip(refs, order)]
    return ops, refs


def test_terminal_measurement_detection():
    from qgate_b200.simulator.simulator import Simulator
    from qgate_b200 import model
    from qgate_b200.preprocess import Preprocessor
    ops, refs = _terminal_circuit()
    flat = Preprocessor(circuit_prep='dynamic').preprocess(model.GateList(ops))
    prefix, measures = Simulator._terminal_measurements(flat)
    assert len(measures) == 7 and not any(isinstance(op, model.Measure) for op in prefix)
    q = S.new_qregs(2)
    r = S.new_references(2)
    mid = [S.H(q[0]), S.measure(r[0], q[0]), S.H(q[1]), S.measure(r[1], q[1])]
    assert Simulator._terminal_measurements(Preprocessor().preprocess(model.GateList(mid))) is None
    cond = [S.H(q[0]), S.measure(r[0], q[0]), S.if_(r[0], 1, S.X(q[1])), S.measure(r[1], q[1])]
    assert Simulator._terminal_measurements(Preprocessor().preprocess(model.GateList(cond))) is None
    twice = [S.H(q[0]), S.measure(r[0], q[0]), S.measure(r[1], q[0])]
    assert Simulator._terminal_measurements(Preprocessor().preprocess(model.GateList(twice))) is None


@pytest.mark.gpu
@pytest.mark.parametrize('prep', ('dynamic', 'one_static'))
@pytest.mark.parametrize('dtype', (np.float64, np.float32))
def test_native_sample_matches_the_per_shot_loop_shot_for_shot(cuda_runtime, ref_runtime, dtype, prep):
    """4000 shots from ONE run + one pool on the engine = the reference CPU runtime re-running the
    circuit 4000 times (simulator.py:85-119) under the same seed, observation for observation."""
    ops, refs = _terminal_circuit()
    ref_order = [refs[3], refs[0], refs[6], refs[2], refs[5]]            # a subset, reordered
    sim = with_runtime(cuda_runtime, dtype=dtype, circuit_prep=prep)
    api = cuda_runtime.get_api()
    api.stats_reset()
    np.random.seed(99)
    got = sim.sample(ops, ref_order, 4000).intarray
    launches = api.stats()['kernel_launches']
    after = np.random.random_sample()           # the global RNG stream stands where the loop leaves it
    sim.terminate()
    loop = with_runtime(ref_runtime.module, dtype=dtype, circuit_prep=prep)
    np.random.seed(99)
    want = loop.sample(ops, ref_order, 4000).intarray
    assert after == np.random.random_sample()
    loop.terminate()
    assert np.array_equal(got, want)
    assert len(set(got.tolist())) > 8
    assert launches < 400, launches              # not 4000 re-runs
    # and the engine's own per-shot loop (the fallback for mid-circuit measurements) agrees too
    sim = with_runtime(cuda_runtime, dtype=dtype, circuit_prep=prep, native_sample=False)
    np.random.seed(99)
    assert np.array_equal(sim.sample(ops, ref_order, 300).intarray, want[:300])
    sim.terminate()
