"""GPU suite, sharded part on ONE GPU: 2, 4 and 8 ranks share device 0.  Every rank drives the CUDA
engine through the C ABI; the control plane is gloo (NCCL refuses two ranks per device), the
amplitudes move through qgate_b200/csrc/dist.cu — the same peer-pointer exchange kernel the
multi-GPU runs use, over CUDA-IPC mappings of the other ranks' shards (here the "peer" memory is
the same device).  This is the reference's "logical devices on one GPU" test mode
(tests/test_multidevice.py:13-21: device_ids=[0]*8) for the one-process-per-shard design, and it
puts the rank predicates, the lane maps, the exchange index arithmetic for 1, 2 and 3 lanes at
once, sharded pools, collapses and joins into the single-GPU `-m gpu` suite.  Expected values: one
unsharded run of the reference CPU runtime (oracle/_ref) on the same circuit.  The multi-GPU twin
(NCCL control plane, NVLink) is tests/test_dist_nccl.py."""
import os
import sys

import numpy as np
import pytest

from tests.test_dist_gloo import REPO, _build, _expected, _free_port, _observe

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, case, dtype_name, prep, queue, push=True):
    try:
        sys.path.insert(0, REPO)
        os.environ['MASTER_ADDR'] = '127.0.0.1'
        os.environ['MASTER_PORT'] = str(port)
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(0)
        dist.init_process_group('gloo', rank=rank, world_size=world)
        import qgate_b200
        from qgate_b200 import cudaruntime, dist as qdist
        cudaruntime.set_preference(device_ids=[0])
        dtype = np.dtype(dtype_name).type
        runtime = qdist.runtime(cudaruntime, exchange='p2p', push=push, shard_min_lanes=6)
        assert runtime.ctx.on_cuda and not runtime.ctx.comm_cuda
        sim = qgate_b200.simulator.with_runtime(runtime, dtype=dtype, circuit_prep=prep)
        q, ops, refs = _build(case)
        np.random.seed(1234 + rank)       # ranks disagree: the runtime must broadcast the draws
        sim.run(ops)
        out = _observe(sim, q, refs)
        out['stats'] = dict(runtime.ctx.stats)
        out['sharded'] = max(qs.g for qs in sim.qubits.qstates_list)
        out['engine_stats'] = cudaruntime.get_api().stats()
        out['backend'] = cudaruntime.get_api().backend_name
        sim.terminate()
        dist.barrier()
        dist.destroy_process_group()
        queue.put((rank, out))
    except Exception:
        import traceback
        queue.put((rank, traceback.format_exc()))


def _run_world(world, case, dtype_name, prep, push=True):
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    queue = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, case, dtype_name, prep, queue, push))
             for r in range(world)]
    for p in procs:
        p.start()
    import queue as queue_mod
    results = {}
    try:
        for _ in range(world):
            # a rank that raised leaves its peers waiting in a collective: report what arrived
            rank, out = queue.get(timeout=180 if not results else 30)
            results[rank] = out
    except queue_mod.Empty:
        pass
    finally:
        for p in procs:
            p.join(timeout=5)
            if p.is_alive():
                p.kill()
    for rank, out in results.items():
        assert not isinstance(out, str), 'rank {} failed:\n{}'.format(rank, out)
    assert len(results) == world, 'ranks {} never answered (hang)'.format(
        sorted(set(range(world)) - set(results)))
    return results


def _need_gpu():
    try:
        import torch
        if torch.cuda.device_count() < 1:
            pytest.skip('needs a GPU')
    except Exception:
        pytest.skip('needs a GPU')


@pytest.mark.parametrize('push', (True, False), ids=('push', 'inplace'))
@pytest.mark.parametrize('world,case', ((2, 'random'), (2, 'zoo'), (2, 'qft'), (2, 'grover'), (2, 'measure'),
                                        (2, 'mcz'), (4, 'random'), (4, 'grover'), (4, 'zoo'), (8, 'random')))
def test_ranks_sharing_one_gpu_match_reference(world, case, push, ref_runtime):
    _need_gpu()
    if not push and (world, case) not in ((2, 'random'), (4, 'random'), (8, 'random'), (2, 'mcz')):
        pytest.skip('the in-place kernel is covered by the exchange-heavy cases')
    results = _run_world(world, case, 'float64', 'one_static', push)
    want = _expected(case, 'float64', 'one_static', 1234)
    for rank, got in results.items():
        assert got['backend'] == 'cuda-sm_100a'
        assert got['sharded'] == int(np.log2(world)), 'state vector was not sharded'
        assert got['engine_stats']['kernel_launches'] > 0
        for key in ('states', 'slice', 'states_rev'):
            assert np.abs(got[key] - want[key]).max() < 1e-12, (rank, key)
        assert np.abs(got['prob'] - want['prob']).max() < 1e-12
        assert np.abs(got['p0'] - want['p0']).max() < 1e-12
        for key in ('samples', 'samples_hidden', 'samples_empty'):
            assert np.array_equal(got[key], want[key]), (rank, key)
        if 'bits' in want:
            assert np.array_equal(got['bits'], want['bits']), rank
    if case in ('random', 'zoo', 'grover', 'mcz'):
        st = results[0]['stats']
        assert st['exchanges'] > 0, 'no lane exchange was exercised'
        assert st.get('push_exchanges', 0) == (st['exchanges'] if push else 0)
        if world >= 4 and case == 'random':
            assert st['exchange_lanes'] > st['exchanges'], 'no multi-lane exchange was exercised'


def test_ranks_sharing_one_gpu_float32(ref_runtime):
    _need_gpu()
    results = _run_world(2, 'random', 'float32', 'one_static')
    want = _expected('random', 'float32', 'one_static', 1234)
    for rank, got in results.items():
        assert got['states'].dtype == np.complex64
        assert np.abs(got['states'] - want['states']).max() < 1e-5
        assert np.mean(got['samples'] == want['samples']) > 0.99


def test_joins_of_sharded_groups_on_one_gpu(ref_runtime):
    """Dynamic qubit grouping at sharded sizes (see tests/test_dist_gloo.py)."""
    _need_gpu()
    results = _run_world(2, 'joins', 'float64', 'dynamic')
    want = _expected('joins', 'float64', 'dynamic', 1234)
    for rank, got in results.items():
        assert got['sharded'] == 1
        for key in ('states', 'slice', 'states_rev'):
            assert np.abs(got[key] - want[key]).max() < 1e-12, (rank, key)
        assert np.abs(got['p0'] - want['p0']).max() < 1e-12
        assert np.array_equal(got['samples'], want['samples'])
        assert np.array_equal(got['bits'], want['bits'])
        assert got['stats']['sharded_joins'] >= 3 and got['stats']['gathered_operands'] >= 1
