#!/usr/bin/env python
"""One process per GPU (torchrun), the state vector sharded on its high-order qubits: the same script as on one
GPU, after torch.distributed is initialised.  `torchrun --nproc-per-node 4 examples/sharded_qft.py 32`"""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
local_rank = int(os.environ.get('LOCAL_RANK', '0'))
torch.cuda.set_device(local_rank)
dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
import qgate_b200                               # noqa: E402
import qgate_b200.script as S                   # noqa: E402
from qgate_b200 import circuits, cudaruntime    # noqa: E402

cudaruntime.set_preference(device_ids=[local_rank])
n = int(sys.argv[1]) if len(sys.argv) > 1 else 30 + int(np.log2(dist.get_world_size()))
q, ops = circuits.qft(S, n)
sim = qgate_b200.simulator.cuda(dtype=np.float64, circuit_prep=qgate_b200.prefs.one_static)
torch.cuda.synchronize()
dist.barrier()
t0 = time.perf_counter()
sim.run(ops)
torch.cuda.synchronize()
dist.barrier()
dt = time.perf_counter() - t0
p = sim.qubits.calc_probability(q[n - 1])        # every rank gets the same number
if dist.get_rank() == 0:
    print('QFT-{} on {} GPUs: {:.1f} ms, P(q[{}] = 0) = {:.6f}'.format(n, dist.get_world_size(), 1e3 * dt, n - 1, p))
sim.terminate()
dist.barrier()
dist.destroy_process_group()
