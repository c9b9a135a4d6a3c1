#!/usr/bin/env python
"""Sharded state vectors of different sizes one after the other in ONE process (what bench.py's sub-records
do): free device memory after every step.  torchrun --nproc-per-node N tools/dist_mem_sequence.py 31 34"""
import os
import sys
import traceback

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
from qgate_b200 import circuits, cudaruntime  # noqa: E402

cudaruntime.set_preference(device_ids=[local_rank])


def free_gib():
    return torch.cuda.mem_get_info()[0] / 2. ** 30


for n in [int(a) for a in sys.argv[1:]]:
    for it in range(2):
        try:
            q, ops = circuits.qft(S, n)
            sim = qgate_b200.simulator.cuda(dtype=np.float64, circuit_prep=qgate_b200.prefs.one_static)
            sim.run(ops)
            p = sim.qubits.calc_probability(q[n - 1])
            torch.cuda.synchronize()
            before = free_gib()
            sim.terminate()
            del sim
            torch.cuda.synchronize()
            dist.barrier()
            if rank == 0:
                print('QFT-%d run %d ok: P0 %.3f, free %.1f GiB while live, %.1f GiB after terminate' % (n, it, p, before, free_gib()), flush=True)
        except Exception:
            print('rank %d QFT-%d run %d FAILED, free %.1f GiB\n%s' % (rank, n, it, free_gib(), traceback.format_exc()[-1500:]), flush=True)
            raise
dist.barrier()
dist.destroy_process_group()
