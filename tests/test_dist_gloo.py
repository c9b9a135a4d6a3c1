"""CPU suite, part 3: the sharding host logic (qgate_b200/dist.py) on world_size 2 and 4 with the
gloo backend.  Each rank drives the reference CPU runtime (oracle/_ref) as its LOCAL engine, so
what is under test is exactly the host side of the multi-GPU path: lane maps, rank predicates for
global controls, diagonal gates on global lanes, the victim choice, the exchange index arithmetic
(collective engine), sharded probability / collapse / readout / sampling pools.  The expected
values come from the same circuit on one unsharded reference runtime."""
import os
import socket
import sys

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _build(case):
    import qgate_b200.script as S
    from qgate_b200 import circuits
    refs = None
    if case == 'random':
        q, ops = circuits.random_u3_cx(S, 9, 6, seed=3)
    elif case == 'zoo':
        q, ops = circuits.mixed_gate_zoo(S, 8, 250, seed=12)
    elif case == 'qft':
        q, ops = circuits.qft(S, 9)
    elif case == 'grover':
        q, ops = circuits.grover(S, 8, 2, 0x5A)
    elif case == 'mcz':
        # diagonal gate on a global lane whose controls cover EVERY local lane, and the same with
        # a non-diagonal target; then enough dense gates to force exchanges both ways
        q = S.new_qregs(7)
        ops = [S.H(qr) for qr in q] + [S.ctrl(q[:-1]).Z(q[-1]), S.ctrl(q[:-1]).U1(0.3)(q[-1]),
                                       S.ctrl(q[1:]).Rz(0.7)(q[0]), S.ctrl(q[:-1]).X(q[-1])]
        ops += [S.Ry(0.2 * (i + 1))(qr) for i, qr in enumerate(q)] + [S.ctrl(q[1:]).Z(q[0])]
        ops += circuits.random_u3_cx(S, 7, 3, seed=5, qregs=q)[1]
    elif case == 'joins':
        # two registers grow independently past the sharding threshold (6 lanes), keep taking in
        # fresh qubits, then get entangled with each other
        a, b = S.new_qregs(7), S.new_qregs(7)
        refs = S.new_references(2)
        ops = []
        for reg, seed in ((a, 1), (b, 2)):
            ops += [S.H(reg[0])] + [S.ctrl(reg[i]).X(reg[i + 1]) for i in range(5)]
            ops += circuits.random_u3_cx(S, 6, 2, seed=seed, qregs=reg[:6])[1]
            ops += [S.Ry(0.4)(reg[6]), S.ctrl(reg[6]).Rx(0.9)(reg[2])]          # sharded + 1 qubit
        ops += [S.ctrl(a[1]).X(b[3]), S.ctrl(b[6]).Ry(1.1)(a[0])]                # sharded + sharded
        ops += circuits.random_u3_cx(S, 14, 2, seed=3, qregs=a + b)[1]
        ops += [S.measure(refs[0], a[3]), S.H(a[3]), S.ctrl(a[3]).X(b[0]), S.measure(refs[1], b[5])]
        q = a + b
    elif case == 'measure':
        q, ops = circuits.random_u3_cx(S, 8, 3, seed=8)
        refs = S.new_references(4)
        ops = ops + [S.measure(refs[0], q[7]), S.measure(refs[1], q[0]), S.H(q[7]),
                     S.ctrl(q[7]).X(q[3]), S.measure(refs[2], q[6]), S.reset(q[6]),
                     S.H(q[6]), S.measure(refs[3], q[7])]
    else:
        raise ValueError(case)
    return q, ops, refs


def _observe(sim, q, refs, dist_ctx=None):
    import qgate_b200.script as S
    sim.qubits.set_ordering(q)
    out = {}
    out['states'] = sim.qubits.states[:]
    out['prob'] = sim.qubits.prob[:]
    out['slice'] = sim.qubits.states[5:400:7]
    out['p0'] = np.array([sim.qubits.calc_probability(qr) for qr in q])
    rnd = np.random.RandomState(9).random_sample(4000)
    out['samples'] = sim.qubits.create_sampling_pool(q).sample(4000, rnd).intarray
    out['samples_hidden'] = sim.qubits.create_sampling_pool(q[1::2]).sample(4000, rnd).intarray
    extra = S.new_qregs(1)
    out['samples_empty'] = sim.qubits.create_sampling_pool(
        [q[2], extra[0]] + list(reversed(q[3:]))).sample(4000, rnd).intarray
    sim.qubits.set_ordering(list(reversed(q)))
    out['states_rev'] = sim.qubits.states[:]
    if refs is not None:
        out['bits'] = np.array(sim.values.get(refs), np.int64)
    return out


def _worker(rank, world, port, case, dtype_name, prep, queue):
    try:
        sys.path.insert(0, REPO)
        os.environ['MASTER_ADDR'] = '127.0.0.1'
        os.environ['MASTER_PORT'] = str(port)
        os.environ['QGATE_NUM_WORKERS'] = '1'
        # world 4 streams the unblocked gates to the local engine every 16 submissions (dist.py
        # _submitted: 2048 by default, more than these circuits hold), world 2 / 8 run them at the flush
        os.environ['QGB_DIST_STREAM'] = '16' if world == 4 else '2048'
        import torch
        import torch.distributed as dist
        torch.set_num_threads(1)
        dist.init_process_group('gloo', rank=rank, world_size=world)
        import qgate_b200
        from qgate_b200 import dist as qdist
        from oracle import ref_runtime
        dtype = np.dtype(dtype_name).type
        runtime = qdist.runtime(ref_runtime.module, shard_min_lanes=6)
        sim = qgate_b200.simulator.with_runtime(runtime, dtype=dtype, circuit_prep=prep)
        q, ops, refs = _build(case)
        np.random.seed(1234 + rank)       # ranks disagree: the runtime must broadcast the draws
        sim.run(ops)
        out = _observe(sim, q, refs)
        out['stats'] = dict(runtime.ctx.stats)
        out['sharded'] = max(qs.g for qs in sim.qubits.qstates_list)
        sim.terminate()
        dist.barrier()
        dist.destroy_process_group()
        queue.put((rank, out))
    except Exception:
        import traceback
        queue.put((rank, traceback.format_exc()))


def _run_world(world, case, dtype_name, prep):
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    queue = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, case, dtype_name, prep, queue))
             for r in range(world)]
    for p in procs:
        p.start()
    results = {}
    for _ in range(world):
        rank, out = queue.get(timeout=240)
        results[rank] = out
    for p in procs:
        p.join(timeout=60)
    for rank, out in results.items():
        assert not isinstance(out, str), 'rank {} failed:\n{}'.format(rank, out)
    return results


def _expected(case, dtype_name, prep, seed_rank0):
    import qgate_b200
    from oracle import ref_runtime
    dtype = np.dtype(dtype_name).type
    sim = qgate_b200.simulator.with_runtime(ref_runtime.module, dtype=dtype, circuit_prep=prep)
    q, ops, refs = _build(case)
    np.random.seed(seed_rank0)
    sim.run(ops)
    return _observe(sim, q, refs)


@pytest.mark.parametrize('world', (2, 4))
@pytest.mark.parametrize('case', ('random', 'zoo', 'qft', 'grover', 'measure', 'mcz'))
def test_sharded_matches_single(world, case, ref_runtime):
    prep = 'one_static'
    results = _run_world(world, case, 'float64', prep)
    want = _expected(case, 'float64', prep, 1234)
    for rank, got in results.items():
        assert got['sharded'] == int(np.log2(world)), 'state vector was not sharded'
        for key in ('states', 'slice', 'states_rev'):
            assert np.abs(got[key] - want[key]).max() < 1e-12, (rank, key)
        assert np.abs(got['prob'] - want['prob']).max() < 1e-12
        assert np.abs(got['p0'] - want['p0']).max() < 1e-12
        for key in ('samples', 'samples_hidden', 'samples_empty'):
            assert np.array_equal(got[key], want[key]), (rank, key)
        if 'bits' in want:
            assert np.array_equal(got['bits'], want['bits']), rank
    # every rank saw the same thing
    for key in ('states', 'samples'):
        assert np.array_equal(results[0][key], results[world - 1][key])
    if case in ('random', 'zoo', 'grover', 'mcz'):
        assert results[0]['stats']['exchanges'] > 0, 'no lane exchange was exercised'


def test_sharded_float32_and_dynamic_prep(ref_runtime):
    results = _run_world(2, 'random', 'float32', 'one_static')
    want = _expected('random', 'float32', 'one_static', 1234)
    for rank, got in results.items():
        assert got['states'].dtype == np.complex64
        assert np.abs(got['states'] - want['states']).max() < 1e-6
        assert np.mean(got['samples'] == want['samples']) > 0.99
    # dynamic grouping keeps the groups below the sharding threshold until the joins: the
    # replicated path and the join into a sharded state are both on this route
    results = _run_world(2, 'qft', 'float64', 'static')
    want = _expected('qft', 'float64', 'static', 1234)
    for rank, got in results.items():
        assert np.abs(got['states'] - want['states']).max() < 1e-12


@pytest.mark.parametrize('world', (2, 4))
def test_joins_of_sharded_groups(world, ref_runtime):
    """Dynamic qubit grouping at sharded sizes: a sharded group absorbs single qubits (one sharded
    operand, no traffic), two sharded groups join (the smaller is gathered into a replica), and a
    measurement then separates a lane again."""
    results = _run_world(world, 'joins', 'float64', 'dynamic')
    want = _expected('joins', 'float64', 'dynamic', 1234)
    for rank, got in results.items():
        assert got['sharded'] == int(np.log2(world))
        for key in ('states', 'slice', 'states_rev'):
            assert np.abs(got[key] - want[key]).max() < 1e-12, (rank, key)
        assert np.abs(got['p0'] - want['p0']).max() < 1e-12
        assert np.array_equal(got['samples'], want['samples'])
        assert np.array_equal(got['bits'], want['bits'])
        assert got['stats']['sharded_joins'] >= 3 and got['stats']['gathered_operands'] >= 1, got['stats']
        assert got['stats']['sharded_joins'] >= 3 and got['stats']['gathered_operands'] >= 1, got['stats']


def test_world_8_three_global_lanes(ref_runtime):
    """g = 3 (the 8-GPU layout): 3-lane exchanges, rank predicates on three global lanes, pools and
    readout over eight shards."""
    results = _run_world(8, 'random', 'float64', 'one_static')
    want = _expected('random', 'float64', 'one_static', 1234)
    for rank, got in results.items():
        assert got['sharded'] == 3
        for key in ('states', 'slice', 'states_rev'):
            assert np.abs(got[key] - want[key]).max() < 1e-12, (rank, key)
        assert np.abs(got['p0'] - want['p0']).max() < 1e-12
        for key in ('samples', 'samples_hidden', 'samples_empty'):
            assert np.array_equal(got[key], want[key]), (rank, key)
    assert results[0]['stats']['exchanges'] > 0
    assert results[0]['stats']['exchange_lanes'] >= results[0]['stats']['exchanges']
