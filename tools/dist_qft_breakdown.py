#!/usr/bin/env python
"""Wall-clock breakdown of a sharded QFT run (torchrun, one rank per GPU): peer mapping, exchanges,
host-side queue handling, passes.  Usage: torchrun --nproc-per-node N tools/dist_qft_breakdown.py [qubits] [repeat]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

rank = int(os.environ.get('RANK', '0'))
local_rank = int(os.environ.get('LOCAL_RANK', '0'))
torch.cuda.set_device(local_rank)
dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
import qgate_b200  # noqa: E402
import qgate_b200.script as S  # noqa: E402
from qgate_b200 import circuits, cudaruntime, dist as D  # noqa: E402

cudaruntime.set_preference(device_ids=[local_rank])
n = int(sys.argv[1]) if len(sys.argv) > 1 else 31
repeat = int(sys.argv[2]) if len(sys.argv) > 2 else 3
acc = {}


def timed(cls, name):
    orig = getattr(cls, name)

    def wrapper(*a, **k):
        t0 = time.perf_counter()
        try:
            return orig(*a, **k)
        finally:
            acc[name] = acc.get(name, 0.) + time.perf_counter() - t0
    setattr(cls, name, wrapper)


for fn in ('_ensure_peers', 'exchange', '_run_pending', '_swap_in', '_apply_unblocked', 'synchronize', 'initialize_qubit_states', 'reset_qubit_states'):
    timed(D.DistQubitProcessor, fn)
api = cudaruntime.get_api()
q, ops = circuits.qft(S, n)
for it in range(repeat):
    acc.clear()
    api.stats_reset()
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    sim = qgate_b200.simulator.cuda(dtype=np.float64, circuit_prep=qgate_b200.prefs.one_static)
    sim.run(ops)
    torch.cuda.synchronize()
    dist.barrier()
    t1 = time.perf_counter()
    st = api.stats()
    if rank == 0:
        print('QFT-%d on %d GPUs, run %d: total %.1f ms | %s | passes %d fans %d' % (
            n, dist.get_world_size(), it, 1e3 * (t1 - t0),
            ' '.join('%s %.1f' % (k, 1e3 * v) for k, v in sorted(acc.items())), st['tile_passes'], st['fan_ops']), flush=True)
    sim.terminate()
    del sim
dist.barrier()
dist.destroy_process_group()
