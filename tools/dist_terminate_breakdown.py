#!/usr/bin/env python
"""Where sim.terminate() of a sharded simulator spends its time (IPC unmapping vs returning blocks to the pool).
torchrun --nproc-per-node N tools/dist_terminate_breakdown.py [qubits] [repeat]"""
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
from qgate_b200 import circuits, cudaruntime, dist as D, native  # noqa: E402

cudaruntime.set_preference(device_ids=[local_rank])
n = int(sys.argv[1]) if len(sys.argv) > 1 else 31
repeat = int(sys.argv[2]) if len(sys.argv) > 2 else 3
acc = {}


def timed(cls, name, label=None):
    orig = getattr(cls, name)
    label = label or name

    def wrapper(*a, **k):
        t0 = time.perf_counter()
        try:
            return orig(*a, **k)
        finally:
            acc[label] = acc.get(label, 0.) + time.perf_counter() - t0
    setattr(cls, name, wrapper)


timed(D.DistQubitStates, '_close_peers')
timed(native.NativeQubitStates, 'delete', 'local.delete')
api = cudaruntime.get_api()
q, ops = circuits.random_u3_cx(S, n, 12, seed=3)
for it in range(repeat):
    sim = qgate_b200.simulator.cuda(dtype=np.float64, circuit_prep=qgate_b200.prefs.one_static)
    sim.run(ops)
    p = sim.qubits.calc_probability(q[n - 1])
    torch.cuda.synchronize()
    dist.barrier()
    acc.clear()
    free0 = torch.cuda.mem_get_info()[0] / 2. ** 30
    t0 = time.perf_counter()
    sim.terminate()
    del sim
    t1 = time.perf_counter()
    if rank == 0:
        print('run %d: terminate %.1f ms | %s | free %.1f -> %.1f GiB' % (
            it, 1e3 * (t1 - t0), ' '.join('%s %.1f' % (k, 1e3 * v) for k, v in sorted(acc.items())), free0,
            torch.cuda.mem_get_info()[0] / 2. ** 30), flush=True)
dist.barrier()
dist.destroy_process_group()
