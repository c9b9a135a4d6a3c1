"""Synthetic circuit generators for the BASELINE.json configs.

Each generator takes the script namespace `S` it should build with — `qgate_b200.script`
or, when generating golden vectors from the real reference, `qgate.script` — so the very
same gate sequence reaches both implementations.  All randomness is seeded.
"""
import math

import numpy as np


def qft(S, n, qregs=None, prepare=True):
    """QFT of examples/quantum_fourier_transform.py:36-50 generalised to n qubits:
    X(q0), X(q2), then for i: H(q[i]); ctrl(q[j]).U1(pi / 2^(j-i))(q[i]) for j > i.
    n + n(n-1)/2 (+2) gates."""
    q = S.new_qregs(n) if qregs is None else qregs
    ops = []
    if prepare:
        ops += [S.X(q[0])]
        if n > 2:
            ops += [S.X(q[2])]
    for i in range(n):
        ops.append(S.H(q[i]))
        for j in range(i + 1, n):
            ops.append(S.ctrl(q[j]).U1(math.pi / float(1 << (j - i)))(q[i]))
    return q, ops


def qft_textbook_order(S, n, qregs=None):
    """Same gates in the order of examples/quantum_fourier_transform.py (controlled phases of
    one control qubit grouped before its H): H(q0); cu1(q1->q0); H(q1); cu1(q2->q0); ..."""
    q = S.new_qregs(n) if qregs is None else qregs
    ops = [S.X(q[0])] + ([S.X(q[2])] if n > 2 else [])
    for j in range(n):
        for i in range(j):
            ops.append(S.ctrl(q[j]).U1(math.pi / float(1 << (j - i)))(q[i]))
        ops.append(S.H(q[j]))
    return q, ops


def random_u3_cx(S, n, depth, seed=1234, qregs=None):
    """Config 2: per layer d, U3 with uniform angles on every qubit, then a CX ladder
    ctrl(q[i]).X(q[i+1]) for i in range(d % 2, n - 1, 2)."""
    rng = np.random.RandomState(seed)
    q = S.new_qregs(n) if qregs is None else qregs
    ops = []
    for d in range(depth):
        for i in range(n):
            theta, phi, lam = rng.uniform(0., 2. * math.pi, 3)
            ops.append(S.U3(float(theta), float(phi), float(lam))(q[i]))
        for i in range(d % 2, n - 1, 2):
            ops.append(S.ctrl(q[i]).X(q[i + 1]))
    return q, ops


def random_circuit_gate_count(n, depth):
    return sum(n + len(range(d % 2, n - 1, 2)) for d in range(depth))


def grover(S, n, iterations, marked, qregs=None):
    """Config 3: H on all; `iterations` x (oracle marking |marked>, diffusion).
    oracle   = [X on zero bits of marked] ctrl(q[:-1]).Z(q[-1]) [X ...]
    diffusion= H^n X^n ctrl(q[:-1]).Z(q[-1]) X^n H^n"""
    q = S.new_qregs(n) if qregs is None else qregs
    ops = [S.H(qr) for qr in q]
    zero_bits = [q[i] for i in range(n) if not (marked >> i) & 1]
    for _ in range(iterations):
        ops += [S.X(qr) for qr in zero_bits]
        ops.append(S.ctrl(q[:-1]).Z(q[-1]))
        ops += [S.X(qr) for qr in zero_bits]
        ops += [S.H(qr) for qr in q]
        ops += [S.X(qr) for qr in q]
        ops.append(S.ctrl(q[:-1]).Z(q[-1]))
        ops += [S.X(qr) for qr in q]
        ops += [S.H(qr) for qr in q]
    return q, ops


def phase_estimation(S, n_bits, v_in):
    """Config 4 (examples/phase_estimation.py:8-65): H on the counting bits,
    ctrl(bit_i).Expii(2 pi v_in 2^i)(target), then the inverse QFT = the QFT gate list
    reversed with every gate adjointed.  The target qreg is created first, like the example."""
    target = S.new_qreg()
    bits = S.new_qregs(n_bits)
    ops = [S.H(b) for b in bits]
    for i, b in enumerate(bits):
        ops.append(S.ctrl(b).Expii(2. * math.pi * v_in * float(1 << i))(target))
    inverse = []
    for i in range(n_bits):
        inverse.append(S.H.Adj(bits[i]))
        for k, cbit in enumerate(bits[i + 1:]):
            inverse.append(S.ctrl(cbit).U1(math.pi / float(1 << (k + 1))).Adj(bits[i]))
    inverse.reverse()
    return bits, target, ops + inverse


def ghz_ladder(S, n, qregs=None):
    q = S.new_qregs(n) if qregs is None else qregs
    return q, [S.H(q[0])] + [S.ctrl(q[i]).X(q[i + 1]) for i in range(n - 1)]


def mixed_gate_zoo(S, n, n_gates, seed, qregs=None):
    """Every gate type x {plain, adjoint} x {0, 1, 2 controls} in a seeded random order;
    used by parity tests to exercise all matrix factories and control paths."""
    rng = np.random.RandomState(seed)
    q = S.new_qregs(n) if qregs is None else qregs
    const = ['I', 'H', 'S', 'T', 'X', 'Y', 'Z', 'SH']
    param = {'Rx': 1, 'Ry': 1, 'Rz': 1, 'U1': 1, 'U2': 2, 'U3': 3, 'Expii': 1, 'Expiz': 1}
    names = const + sorted(param)
    ops = [S.H(qr) for qr in q]
    for _ in range(n_gates):
        name = names[rng.randint(len(names))]
        n_ctrl = int(rng.randint(0, min(3, n)))
        lanes = rng.permutation(n)[:n_ctrl + 1]
        target, ctrls = q[int(lanes[0])], [q[int(l)] for l in lanes[1:]]
        holder = S.ctrl(ctrls) if n_ctrl else S
        factory = getattr(holder, name)
        if name in param:
            factory = factory(*[float(v) for v in rng.uniform(-math.pi, math.pi, param[name])])
        if rng.randint(2):
            factory = factory.Adj
        ops.append(factory(target))
    return q, ops
