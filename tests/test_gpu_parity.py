"""GPU suite: the sm_100a engine, driven through the C ABI by the same host stack, against
(1) the golden vectors recorded from the real reference (tests/golden/), (2) the unmodified
reference CPU runtime (oracle/_ref) run on the same seeded inputs, and (3) size-independent
properties at BASELINE-scale qubit counts.

Tolerances are BASELINE.json's: amplitudes within 1e-12 (complex128) / 1e-5 (complex64)
relative to max|a|; measurement outcomes and complex128 sampled indices bit-exact for
identical random draws."""
import math

import numpy as np
import pytest

import qgate_b200
import qgate_b200.script as S
from qgate_b200 import circuits
from tests import cases

pytestmark = pytest.mark.gpu

PREPS = ('dynamic', 'one_static')
RTS = (('cpu64', np.float64), ('cpu32', np.float32))
PROB_TOL = {np.float64: 1e-12, np.float32: 2e-6}


@pytest.fixture(autouse=True)
def default_options(cuda_runtime):
    api = cuda_runtime.get_api()
    for name, value in (('fuse', 1), ('merge', 1), ('tile_lanes_fp64', 10), ('tile_lanes_fp32', 11),
                        ('low_lanes_fp64', 0), ('low_lanes_fp32', 0), ('max_gates_per_pass', 112), ('max_cost', 0),
                        ('tma_buffers', 0), ('reg_bits_fp64', 4), ('ctas_per_sm', 0), ('tma_ws', 0),
                        ('warp_local', 0), ('shear_fp64', 1), ('shear_fp32', 1), ('exact', 0), ('l2_hint', 0),
                        ('debug_pass_mode', 0)):
        api.set_option(name, value)
    yield


@pytest.mark.parametrize('name', sorted(cases.CIRCUITS))
@pytest.mark.parametrize('prep', PREPS)
@pytest.mark.parametrize('rt,dtype', RTS)
def test_states_prob_p0_vs_golden(golden, cuda_runtime, name, prep, rt, dtype):
    sim, q = cases.run_circuit(cuda_runtime, name, dtype, prep)
    key = '{}/{}/{}'.format(name, prep, rt)
    states = sim.qubits.states[:]
    assert states.dtype == golden[key + '/states'].dtype
    assert cases.rel_err(states, golden[key + '/states']) < cases.TOL[dtype]
    if dtype is np.float32:  # also against the complex128 reference run
        assert cases.rel_err(states, golden['{}/{}/cpu64/states'.format(name, prep)]) < 1e-5
    assert np.abs(sim.qubits.prob[:] - golden[key + '/prob']).max() < PROB_TOL[dtype]
    p0 = np.array([sim.qubits.calc_probability(qr) for qr in q])
    assert np.abs(p0 - golden[key + '/p0']).max() < (1e-12 if dtype is np.float64 else 1e-5)
    assert cuda_runtime.get_api().backend_name == 'cuda-sm_100a'


@pytest.mark.parametrize('name', ('rand6x10', 'zoo7', 'grover8', 'ghz11'))
@pytest.mark.parametrize('prep', PREPS)
@pytest.mark.parametrize('rt,dtype', RTS)
def test_measure_all_bits_exact(golden, cuda_runtime, name, prep, rt, dtype):
    sim = cases.make_sim(cuda_runtime, dtype, prep)
    q, ops = cases.CIRCUITS[name]()
    refs = S.new_references(len(q))
    np.random.seed(11)
    sim.run(ops + [S.measure(r, qr) for r, qr in zip(refs, q)])
    sim.qubits.set_ordering(q)
    key = 'measure/{}/{}/{}'.format(name, prep, rt)
    assert np.array_equal(np.array(sim.values.get(refs), np.int64), golden[key + '/bits'])
    assert cases.rel_err(sim.qubits.states[:], golden[key + '/states']) < cases.TOL[dtype] * 10


@pytest.mark.parametrize('prep', PREPS)
@pytest.mark.parametrize('rt,dtype', RTS)
def test_midcircuit_measure_reset_if(golden, cuda_runtime, prep, rt, dtype):
    sim = cases.make_sim(cuda_runtime, dtype, prep)
    q, refs, ops = cases.midcircuit_ops()
    np.random.seed(5)
    sim.run(ops)
    sim.qubits.set_ordering(q)
    key = 'measure/midcircuit/{}/{}'.format(prep, rt)
    assert np.array_equal(np.array(sim.values.get(refs), np.int64), golden[key + '/bits'])
    assert cases.rel_err(sim.qubits.states[:], golden[key + '/states']) < cases.TOL[dtype] * 10


class Probe:
    def __init__(self, prob, empty_lanes, qreg_ordering):
        self.prob = prob


@pytest.mark.parametrize('name', ('rand10x20', 'zoo9', 'grover8'))
@pytest.mark.parametrize('prep', PREPS)
@pytest.mark.parametrize('rt,dtype', RTS)
def test_sampling_pool_indices(golden, cuda_runtime, name, prep, rt, dtype):
    rnd = np.random.RandomState(7).random_sample(20000)
    sim = cases.make_sim(cuda_runtime, dtype, prep)
    q, ops = cases.CIRCUITS[name]()
    empty = S.new_qregs(2)
    sim.run(ops)
    key = 'sampling/{}/{}/{}'.format(name, prep, rt)
    orderings = {'full': q, 'hidden': q[1::2],
                 'empty': [q[3], empty[0], q[0], q[5], empty[1], q[1]],
                 'reversed': list(reversed(q))}
    for tag, ordering in orderings.items():
        got = sim.qubits.create_sampling_pool(ordering).sample(20000, rnd).intarray
        want = golden[key + '/' + tag]
        if dtype is np.float64:
            assert np.array_equal(got, want), tag
        else:
            # complex64: the reference accumulates its cumulative sum in float32 (error ~1e-6,
            # dependent on its worker count); the engine accumulates in float64.  Indices agree
            # except for draws that fall inside that band.
            assert np.mean(got == want) > 0.995, (tag, np.mean(got == want))
    prob = sim.qubits.create_sampling_pool(q[1::2], Probe).prob
    assert prob.dtype == np.dtype(dtype)
    assert np.allclose(prob, golden[key + '/prob_hidden'], rtol=1e-12 if dtype is np.float64 else 1e-4,
                       atol=1e-15 if dtype is np.float64 else 1e-8)


def reference_cumprob32(prob32, n_workers):
    """The reference's complex64 cumulative array, restated: every worker runs a SEQUENTIAL float32
    running sum over its span (spans of ceil(N / W) rounded up to 16), the span totals are
    prefix-summed in float32, then every entry gets its span offset added and is multiplied by
    1 / total (CPUSamplingPool.cpp:13-47, Parallel.h:12-23; one worker below 2^16 entries,
    Parallel.cpp:60-67)."""
    n = len(prob32)
    w = n_workers if n > (1 << 16) else 1
    span = -(-n // w)
    span = -(-span // 16) * 16
    cum = np.empty(n, np.float32)
    partial = np.zeros(w, np.float32)
    for t in range(w):
        b, e = min(span * t, n), min(span * (t + 1), n)
        if e > b:
            cum[b:e] = np.cumsum(prob32[b:e], dtype=np.float32)
            partial[t] = cum[e - 1]
    partial = np.cumsum(partial, dtype=np.float32)
    norm = np.float32(1.) / partial[-1]
    for t in range(w):
        b, e = min(span * t, n), min(span * (t + 1), n)
        if t > 0:
            cum[b:e] += partial[t - 1]
        cum[b:e] *= norm
    return cum


def assert_fp32_samples_within_reference_band(got, want, ref_prob32, rnd):
    """complex64 pools.  The reference accumulates its cumulative array in float32, sequentially
    per worker: at 2^20 bins its rounding error is tens of bin widths and depends on the worker
    count.  The engine scans in float64, so its indices are the exact upper bounds and cannot be
    identical to the reference's.  What is checked: (1) the restated float32 algorithm reproduces
    the reference's indices exactly (so the error model is the right one), (2) the engine's
    indices are the exact float64 answer, (3) every disagreement lies inside the reference's own
    float32 accumulation error."""
    import os
    r32 = rnd.astype(np.float32)
    last = len(ref_prob32) - 1
    cum32 = None
    for workers in sorted({len(os.sched_getaffinity(0)), os.cpu_count(), 1}, reverse=True):
        cand = reference_cumprob32(ref_prob32, workers)
        if np.array_equal(np.minimum(np.searchsorted(cand, r32, side='right'), last), want):
            cum32 = cand.astype(np.float64)
            break
    assert cum32 is not None, 'restated float32 pool does not reproduce the reference indices'
    cum64 = np.cumsum(ref_prob32.astype(np.float64))
    cum64 *= 1. / cum64[-1]
    r = r32.astype(np.float64)
    exact = np.minimum(np.searchsorted(cum64, r, side='right'), last)
    # the engine's probabilities differ from the reference's by complex64 amplitude rounding only
    assert np.mean(got == exact) > 0.98
    # a draw is answered differently only when it falls between the float64 and the float32 value
    # of one cumulative entry: r is then within the float32 accumulation error of that entry
    band = 2. * np.abs(cum32 - cum64).max() + 4. * np.finfo(np.float32).eps
    mism = np.flatnonzero(got != want)
    first = np.minimum(got[mism], want[mism])
    assert np.abs(cum64[first] - r[mism]).max(initial=0.) <= band, band


@pytest.mark.parametrize('dtype', (np.float64, np.float32))
def test_sampling_is_searchsorted_of_own_prob_array(cuda_runtime, dtype):
    """The pool's documented algorithm: inclusive scan in float64 of the marginal probability
    vector, scaled by 1/total, then upper_bound (CPUSamplingPool.cpp:8-81 in float64)."""
    sim, q = cases.run_circuit(cuda_runtime, 'rand13x8', dtype, 'one_static')
    rnd = np.random.RandomState(21).random_sample(50000)
    got = sim.qubits.create_sampling_pool(q).sample(50000, rnd).intarray
    prob = sim.qubits.create_sampling_pool(q, Probe).prob.astype(np.float64)
    cum = np.cumsum(prob)
    cum *= 1. / cum[-1]
    r = rnd if dtype is np.float64 else rnd.astype(np.float32).astype(np.float64)
    want = np.minimum(np.searchsorted(cum, r, side='right'), len(cum) - 1)
    mismatch = np.flatnonzero(got != want)
    # a different (tiled) summation order may move a boundary by a few ulps
    assert len(mismatch) <= 2
    for i in mismatch:
        assert abs(int(got[i]) - int(want[i])) == 1
        assert abs(cum[min(got[i], want[i])] - r[i]) < 1e-13


@pytest.mark.parametrize('rt,dtype', RTS)
def test_get_states_slices(golden, cuda_runtime, rt, dtype):
    sim, q = cases.run_circuit(cuda_runtime, 'rand10x20', dtype, 'dynamic')
    extra = S.new_qregs(1)
    tol = cases.TOL[dtype]
    for idx, k in enumerate(cases.SLICE_KEYS):
        want = golden['slices/{}/{}/states'.format(rt, idx)]
        got = sim.qubits.states[slice(*k)]
        assert got.shape == want.shape, k
        if want.size:
            assert np.abs(got - want).max() < tol, k
        wantp = golden['slices/{}/{}/prob'.format(rt, idx)]
        gotp = sim.qubits.prob[slice(*k)]
        assert gotp.shape == wantp.shape
        if wantp.size:
            assert np.abs(gotp - wantp).max() < tol, k
    sim.qubits.set_ordering(q[:4] + extra + q[4:])
    want = golden['slices/{}/empty_lane/states'.format(rt)]
    assert np.abs(sim.qubits.states[::5] - want).max() < tol
    sim.qubits.set_ordering(q)
    with pytest.raises(RuntimeError):
        sim.qubits.states[1 << 10]


@pytest.mark.parametrize('rt,dtype', RTS)
def test_phase_estimation_hidden_target(golden, cuda_runtime, rt, dtype):
    sim = cases.make_sim(cuda_runtime, dtype, 'dynamic')
    bits, target, ops = circuits.phase_estimation(S, 8, 0.1)
    sim.run(ops)
    sim.qubits.set_ordering(bits + [target])
    assert cases.rel_err(sim.qubits.states[:], golden['pe8/{}/states'.format(rt)]) < cases.TOL[dtype]
    rnd = np.random.RandomState(3).random_sample(4096)
    got = sim.qubits.create_sampling_pool(bits).sample(4096, rnd).intarray
    if dtype is np.float64:
        assert np.array_equal(got, golden['pe8/{}/samples'.format(rt)])
    else:
        assert np.mean(got == golden['pe8/{}/samples'.format(rt)]) > 0.995


def test_qft20_config0(golden, cuda_runtime):
    """BASELINE.json configs[0]: 20-qubit QFT, complex128."""
    sim = cases.make_sim(cuda_runtime, np.float64, 'one_static')
    q, ops = circuits.qft(S, 20)
    sim.run(ops)
    sim.qubits.set_ordering(q)
    scale = 2. ** -10
    assert np.abs(sim.qubits.states[::4099] - golden['qft20/cpu64/states_stride4099']).max() < 1e-12 * scale
    assert np.abs(sim.qubits.states[:512] - golden['qft20/cpu64/states_head']).max() < 1e-12 * scale
    for qr in q:
        assert abs(sim.qubits.calc_probability(qr) - 0.5) < 1e-12


@pytest.mark.parametrize('dtype', (np.float64, np.float32))
@pytest.mark.parametrize('gen', ('random', 'zoo', 'grover', 'qft'))
def test_against_reference_cpu_runtime_16_to_20_qubits(cuda_runtime, ref_runtime, dtype, gen):
    """Same seeded circuit on the engine and on the unmodified reference CPU runtime."""
    def build():
        if gen == 'random':
            return circuits.random_u3_cx(S, 20, 6, seed=77)
        if gen == 'zoo':
            return circuits.mixed_gate_zoo(S, 16, 600, seed=9)
        if gen == 'grover':
            return circuits.grover(S, 18, 3, 0x2AAAA)
        return circuits.qft(S, 19)
    q, ops = build()
    refs = S.new_references(3)
    tail = [S.measure(refs[0], q[2]), S.measure(refs[1], q[len(q) - 1]), S.H(q[2]),
            S.measure(refs[2], q[2])]
    outs = []
    for rt in (cuda_runtime, ref_runtime.module):
        sim = cases.make_sim(rt, dtype, 'one_static')
        np.random.seed(123)
        sim.run(ops + tail)
        sim.qubits.set_ordering(q)
        p0 = [sim.qubits.calc_probability(qr) for qr in q]
        rnd = np.random.RandomState(4).random_sample(100000)
        samples = sim.qubits.create_sampling_pool(q).sample(100000, rnd).intarray
        prob = sim.qubits.create_sampling_pool(q, Probe).prob
        outs.append((sim.qubits.states[:], sim.values.get(refs), np.array(p0), samples, prob))
        sim.terminate()
    (a, bits_a, p0_a, s_a, prob_a), (b, bits_b, p0_b, s_b, prob_b) = outs
    assert bits_a == bits_b
    if dtype is np.float64:
        assert cases.rel_err(a, b) < cases.TOL[dtype]
    else:
        # the complex64 reference carries its own rounding error (one rounding per gate; the
        # engine composes merged gates in double first): both are held to 1e-5 of the complex128
        # reference, and to each other within the sum of the two errors
        sim = cases.make_sim(ref_runtime.module, np.float64, 'one_static')
        np.random.seed(123)
        sim.run(ops + tail)
        sim.qubits.set_ordering(q)
        exact = sim.qubits.states[:]
        sim.terminate()
        err_ref = cases.rel_err(b, exact)
        assert cases.rel_err(a, exact) < cases.TOL[dtype]
        assert cases.rel_err(a, b) < cases.TOL[dtype] + err_ref
    assert np.abs(p0_a - p0_b).max() < (1e-12 if dtype is np.float64 else 1e-5)
    if dtype is np.float64:
        assert np.array_equal(s_a, s_b)
    else:
        assert_fp32_samples_within_reference_band(s_a, s_b, prob_b, rnd)


@pytest.mark.parametrize('dtype', (np.float64, np.float32))
def test_fused_equals_unfused_and_tile_shapes(cuda_runtime, dtype):
    """The planner may only reorder commuting gates: every tile shape, and the one-kernel-per-gate
    path, must give the same amplitudes."""
    api = cuda_runtime.get_api()
    q, ops = circuits.mixed_gate_zoo(S, 15, 500, seed=31)
    q2, ops2 = circuits.random_u3_cx(S, 15, 12, seed=8, qregs=q)
    ops = ops + ops2

    def run(**options):
        for k, v in options.items():
            api.set_option(k, v)
        sim = cases.make_sim(cuda_runtime, dtype, 'one_static')
        sim.run(ops)
        sim.qubits.set_ordering(q)
        out = sim.qubits.states[:]
        sim.terminate()
        return out

    base = run(fuse=0)
    tol = 1e-13 if dtype is np.float64 else 2e-6
    shapes = [dict(fuse=1, merge=0), dict(fuse=1, merge=1)]
    lanes = 'tile_lanes_fp64' if dtype is np.float64 else 'tile_lanes_fp32'
    low = 'low_lanes_fp64' if dtype is np.float64 else 'low_lanes_fp32'
    k = 3 if dtype is np.float64 else 4
    # every tile shape, 2 and 3 tile buffers, dense gates as direct 2x2 products and as shears
    for shear in (0, 1):
        for buffers in (2, 3):
            for t in (k + 5, k + 6, k + 8, k + 9, k + 10):
                for l in (1, 3, 6):
                    shapes.append({'fuse': 1, 'shear': shear, 'tma_buffers': buffers, lanes: t, low: l})
    shapes.append(dict(fuse=1, shear=1, max_gates_per_pass=3))
    shapes.append(dict(fuse=1, shear=0, max_gates_per_pass=3))
    for l2_hint in (1, 2, 3):
        shapes.append(dict(fuse=1, shear=1, max_gates_per_pass=112, l2_hint=l2_hint))
    for ws_buffers in (2, 3):                        # the warp-specialised variant (producer warp)
        for t in (k + 5, k + 7, k + 8, k + 9):
            shapes.append({'fuse': 1, 'shear': 1, 'l2_hint': 0, 'tma_ws': 1, 'tma_buffers': ws_buffers, lanes: t, low: 0})
    shapes.append({'fuse': 1, 'tma_ws': 0, 'tma_buffers': 0, lanes: k + 8, low: 0})
    for t in (k + 6, k + 7, k + 8, k + 9):           # warp-local stage transitions
        shapes.append({'fuse': 1, 'warp_local': 1, lanes: t, low: 0})
    shapes.append({'fuse': 1, 'warp_local': 0, lanes: k + 6, low: 0})
    if dtype is np.float64:
        for t in (8, 9, 11, 12):
            for shear in (0, 1):
                shapes.append(dict(fuse=1, shear=shear, tma_buffers=2, reg_bits_fp64=3, tile_lanes_fp64=t, low_lanes_fp64=5))
    for options in shapes:
        api.stats_reset()
        options.setdefault('shear_fp64', 1)
        options.setdefault('shear_fp32', 1)
        if 'shear' in options:       # (sets both precisions)
            options['shear_fp64'] = options['shear_fp32'] = options.pop('shear')
        got = run(**options)
        assert cases.rel_err(got, base) < tol, options
        stats = api.stats()
        assert stats['tma_passes'] == stats['tile_passes'] > 0, (options, stats)
        if options['shear_fp64' if dtype is np.float64 else 'shear_fp32']:
            assert stats['shear_ops'] > 10 * stats['direct_ops'], (options, stats)
        else:
            assert stats['shear_ops'] == 0 and stats['direct_ops'] > 0, (options, stats)


@pytest.mark.parametrize('dtype,n', ((np.float64, 27), (np.float32, 28)))
def test_large_state_properties(cuda_runtime, dtype, n):
    """BASELINE-scale checks that need no oracle: norm, per-qubit probabilities and amplitude
    slices of QFT|x> against the closed form, U followed by U^dagger returning |0...0>."""
    sim = cases.make_sim(cuda_runtime, dtype, 'one_static')
    q, ops = circuits.qft(S, n)     # X(q0) X(q2) -> |x=5>, then QFT
    sim.run(ops)
    sim.qubits.set_ordering(q)
    tol = cases.TOL[dtype]
    amp = 2. ** (-n / 2.)
    # QFT without the final swaps on |x>, x = 5 (q0 is the LSB): amplitude of |k> is
    # exp(2 pi i k rev_n(x) / 2^n) / sqrt(2^n); validated against the reference CPU runtime at
    # 9 and 12 qubits (max |diff| 1.4e-17).
    start, step = 12345, 65537
    got = sim.qubits.states[start::step]
    ks = np.arange(start, 1 << n, step, dtype=np.int64)
    xr = (1 << (n - 1)) + (1 << (n - 3))
    phase = (ks * xr) % (1 << n)
    want = amp * np.exp(2j * np.pi * phase.astype(np.float64) / float(1 << n))
    assert np.abs(got - want).max() < tol * amp * (4 if dtype is np.float32 else 1)
    for qr in (q[0], q[5], q[n // 2], q[n - 1]):
        assert abs(sim.qubits.calc_probability(qr) - 0.5) < (1e-12 if dtype is np.float64 else 1e-5)
    # inverse: all gates adjointed in reverse order
    inverse = []
    for i in range(n):
        inverse.append(S.H(q[i]))
        for j in range(i + 1, n):
            inverse.append(S.ctrl(q[j]).U1(-math.pi / float(1 << (j - i)))(q[i]))
    sim.run(list(reversed(inverse)))
    head = sim.qubits.states[:8]
    assert abs(abs(head[5]) - 1.) < (1e-11 if dtype is np.float64 else 2e-5)
    assert np.abs(np.delete(head, 5)).max() < (1e-11 if dtype is np.float64 else 2e-5)
    sim.terminate()


def test_errors_surface_as_exceptions(cuda_runtime):
    api = cuda_runtime.get_api()
    with pytest.raises(ValueError):
        api.call('qgb_qstates_delete', 12345)
    sim = cases.make_sim(cuda_runtime, np.float64, 'dynamic')
    q = S.new_qregs(3)
    sim.run([S.H(q[0]), S.ctrl(q[0]).X(q[1])])
    with pytest.raises(RuntimeError):
        sim.run([S.reset(q[1])])   # not measured (rop_executor.py:45-51)


@pytest.mark.gpu
@pytest.mark.parametrize('dtype', (np.float64, np.float32))
def test_simulator_sample_shot_for_shot(cuda_runtime, ref_runtime, dtype):
    """Simulator.sample (simulator.py:85-119): every shot re-runs the circuit with dynamic qubit
    grouping and one RNG draw per Measure; the observations equal the reference CPU runtime's."""
    q = S.new_qregs(4)
    refs = S.new_references(4)
    ops = [S.H(q[0]), S.ctrl(q[0]).X(q[1]), S.Ry(0.7)(q[2]), S.ctrl(q[1]).Rx(1.1)(q[2]),
           S.measure(refs[0], q[0]), S.measure(refs[1], q[1]), S.if_(refs[1], 1, S.X(q[2])),
           S.ctrl(q[2]).H(q[3]), S.measure(refs[2], q[2]), S.reset(q[2]), S.H(q[2]),
           S.measure(refs[3], q[3])]
    outs = []
    for rt in (cuda_runtime, ref_runtime.module):
        sim = cases.make_sim(rt, dtype, 'dynamic')
        np.random.seed(21)
        outs.append(sim.sample(ops, refs, 200).intarray)
        sim.terminate()
    assert np.array_equal(outs[0], outs[1])
    assert len(set(outs[0].tolist())) > 3
