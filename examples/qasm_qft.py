#!/usr/bin/env python
"""OpenQASM 2.0 text -> circuit (qgate_b200.openqasm) -> the engine: a 20-qubit QFT of |5>, amplitudes against the
closed form, then 10 samples of the register."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qgate_b200
from qgate_b200 import openqasm

n = 20
lines = ['OPENQASM 2.0;', 'include "qelib1.inc";', 'qreg q[{}];'.format(n), 'x q[0];', 'x q[2];']
for t in range(n):
    lines.append('h q[{}];'.format(t))
    lines += ['cu1(pi/{}) q[{}], q[{}];'.format(1 << (c - t), c, t) for c in range(t + 1, n)]
prog = openqasm.load_circuit('\n'.join(lines))
sim = qgate_b200.simulator.cuda(dtype=np.float64)
sim.run(prog.circuit)
sim.qubits.set_ordering(prog.q)
k = np.arange(8)
rev5 = (1 << (n - 1)) + (1 << (n - 3))                     # bit-reversed 5 (this QFT has no final swaps)
want = 2. ** (-n / 2.) * np.exp(2j * np.pi * ((k * rev5) % (1 << n)) / float(1 << n))
print('max |amplitude - closed form| over the first 8:', np.abs(sim.qubits.states[:8] - want).max())
pool = sim.qubits.create_sampling_pool(prog.q)
print('samples:', pool.sample(10).intarray)
sim.terminate()
