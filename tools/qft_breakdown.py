#!/usr/bin/env python
"""Where the wall-clock time of one QFT run goes: front end (preprocess + per-gate dispatch), state
initialisation, the fused passes.  Usage: python tools/qft_breakdown.py [qubits] [repeat]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import qgate_b200  # noqa: E402
import qgate_b200.script as S  # noqa: E402
from qgate_b200 import circuits, cudaruntime, model  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
repeat = int(sys.argv[2]) if len(sys.argv) > 2 else 3
cudaruntime.set_preference(device_ids=[0])
api = cudaruntime.get_api()
for opt in sys.argv[3:]:
    k, v = opt.split('=')
    api.set_option(k, int(v))
q, ops = circuits.qft(S, n)
for it in range(repeat):
    sim = qgate_b200.simulator.cuda(dtype=np.float64, circuit_prep=qgate_b200.prefs.one_static)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    pre = sim.preprocessor.preprocess(model.GateList(ops))
    sim._value_store.sync_refs(sim.preprocessor.get_refset())
    t1 = time.perf_counter()
    head = [op for op in pre if not isinstance(op, model.Gate)]
    gates = [op for op in pre if isinstance(op, model.Gate)]
    sim._execute(head)                   # qreg creation: allocation + |0...0>
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    sim._execute(gates)                  # per-gate dispatch into the engine's queue
    t3 = time.perf_counter()
    api.stats_reset()
    sim._qubits.update_external_layout()
    for qstates in sim._qubits.qstates_list:
        qstates.processor.synchronize()  # flush: planner + passes
    torch.cuda.synchronize()
    t4 = time.perf_counter()
    st = api.stats()
    print('QFT-%d run %d: preprocess %.1f ms | create + reset %.1f ms | dispatch of %d gates %.1f ms | flush (plan + %d passes) %.1f ms | total %.1f ms'
          % (n, it, 1e3 * (t1 - t0), 1e3 * (t2 - t1), len(gates), 1e3 * (t3 - t2), st['tile_passes'], 1e3 * (t4 - t3), 1e3 * (t4 - t0)))
    sim.terminate()
    del sim, qstates, pre, head, gates
