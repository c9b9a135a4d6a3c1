"""GPU suite, sharded part: the same cases as tests/test_dist_gloo.py, but every rank drives the
CUDA engine through the C ABI on its own GPU and the ranks talk over NCCL / NVLink peer memory.
Both exchange engines ('p2p' = qgate_b200/csrc/dist.cu through CUDA-IPC peer pointers, 'collective'
= batched NCCL send/recv) are compared with one unsharded run of the reference CPU runtime
(oracle/_ref) on the same circuit.  Needs >= 2 GPUs (skipped otherwise): run with
`gpurun --gpus 2 -- python -m pytest tests/test_dist_nccl.py -m gpu`."""
import os
import sys

import numpy as np
import pytest

from tests.test_dist_gloo import REPO, _build, _expected, _free_port, _observe

pytestmark = pytest.mark.gpu


def _n_gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def _worker(rank, world, port, case, dtype_name, prep, exchange, queue):
    try:
        sys.path.insert(0, REPO)
        os.environ['MASTER_ADDR'] = '127.0.0.1'
        os.environ['MASTER_PORT'] = str(port)
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(rank)
        dist.init_process_group('nccl', rank=rank, world_size=world,
                                device_id=torch.device('cuda', rank))
        import qgate_b200
        from qgate_b200 import cudaruntime, dist as qdist
        cudaruntime.set_preference(device_ids=[rank])
        dtype = np.dtype(dtype_name).type
        runtime = qdist.runtime(cudaruntime, exchange='p2p' if exchange == 'p2p_inplace' else exchange,
                                push=exchange != 'p2p_inplace', shard_min_lanes=6)
        sim = qgate_b200.simulator.with_runtime(runtime, dtype=dtype, circuit_prep=prep)
        q, ops, refs = _build(case)
        np.random.seed(1234 + rank)       # ranks disagree: the runtime must broadcast the draws
        sim.run(ops)
        out = _observe(sim, q, refs)
        out['stats'] = dict(runtime.ctx.stats)
        out['sharded'] = max(qs.g for qs in sim.qubits.qstates_list)
        out['engine_stats'] = cudaruntime.get_api().stats()
        sim.terminate()
        dist.barrier()
        dist.destroy_process_group()
        queue.put((rank, out))
    except Exception:
        import traceback
        queue.put((rank, traceback.format_exc()))


def _run_world(world, case, dtype_name, prep, exchange):
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    queue = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker,
                         args=(r, world, port, case, dtype_name, prep, exchange, queue))
             for r in range(world)]
    for p in procs:
        p.start()
    import queue as queue_mod
    results = {}
    try:
        for _ in range(world):
            # a rank that raised leaves its peers waiting in a collective: report what arrived
            rank, out = queue.get(timeout=120 if not results else 20)
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


@pytest.mark.parametrize('exchange', ('p2p', 'p2p_inplace', 'collective'))
@pytest.mark.parametrize('case', ('random', 'zoo', 'qft', 'grover', 'measure', 'mcz'))
def test_sharded_cuda_matches_reference(case, exchange, ref_runtime):
    if _n_gpus() < 2:
        pytest.skip('needs >= 2 GPUs')
    world = 4 if _n_gpus() >= 4 and case in ('random', 'grover') else 2
    results = _run_world(world, case, 'float64', 'one_static', exchange)
    want = _expected(case, 'float64', 'one_static', 1234)
    for rank, got in results.items():
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
        assert results[0]['stats']['exchanges'] > 0, 'no lane exchange was exercised'
        pushes = results[0]['stats'].get('push_exchanges', 0)
        assert (pushes == results[0]['stats']['exchanges']) if exchange == 'p2p' else pushes == 0


@pytest.mark.parametrize('exchange', ('p2p', 'collective'))
def test_sharded_cuda_float32(exchange, ref_runtime):
    if _n_gpus() < 2:
        pytest.skip('needs >= 2 GPUs')
    results = _run_world(2, 'random', 'float32', 'one_static', exchange)
    want = _expected('random', 'float32', 'one_static', 1234)
    for rank, got in results.items():
        assert got['states'].dtype == np.complex64
        assert np.abs(got['states'] - want['states']).max() < 1e-5
        assert np.mean(got['samples'] == want['samples']) > 0.99


@pytest.mark.parametrize('exchange', ('p2p', 'collective'))
def test_joins_of_sharded_groups_cuda(exchange, ref_runtime):
    """Dynamic qubit grouping at sharded sizes on GPUs (see tests/test_dist_gloo.py)."""
    if _n_gpus() < 2:
        pytest.skip('needs >= 2 GPUs')
    results = _run_world(2, 'joins', 'float64', 'dynamic', exchange)
    want = _expected('joins', 'float64', 'dynamic', 1234)
    for rank, got in results.items():
        assert got['sharded'] == 1
        for key in ('states', 'slice', 'states_rev'):
            assert np.abs(got[key] - want[key]).max() < 1e-12, (rank, key)
        assert np.abs(got['p0'] - want['p0']).max() < 1e-12
        assert np.array_equal(got['samples'], want['samples'])
        assert np.array_equal(got['bits'], want['bits'])
        assert got['stats']['sharded_joins'] >= 3 and got['stats']['gathered_operands'] >= 1
