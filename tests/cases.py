"""Shared parity cases: every case builds its circuit with qgate_b200.script through
qgate_b200/circuits.py — the same generators tests/golden/make_golden.py fed to the real
reference — and returns what the golden file recorded for it."""
import numpy as np

import qgate_b200
import qgate_b200.script as S
from qgate_b200 import circuits

CIRCUITS = {
    'qft8': lambda: circuits.qft(S, 8),
    'qft10': lambda: circuits.qft(S, 10),
    'qft_tb9': lambda: circuits.qft_textbook_order(S, 9),
    'rand6x10': lambda: circuits.random_u3_cx(S, 6, 10, seed=1234),
    'rand10x20': lambda: circuits.random_u3_cx(S, 10, 20, seed=99),
    'rand13x8': lambda: circuits.random_u3_cx(S, 13, 8, seed=5),
    'zoo5': lambda: circuits.mixed_gate_zoo(S, 5, 120, seed=1),
    'zoo7': lambda: circuits.mixed_gate_zoo(S, 7, 200, seed=2),
    'zoo9': lambda: circuits.mixed_gate_zoo(S, 9, 300, seed=3),
    'grover8': lambda: circuits.grover(S, 8, 2, 0xAA),
    'ghz11': lambda: circuits.ghz_ladder(S, 11),
}

SLICE_KEYS = [(None, None, None), (3, 200, 7), (None, None, -1), (500, 20, -3), (-5, None, 1),
              (10, 11, 1), (0, 1024, 300), (1023, None, -1023), (-1, -1025, -1)]

# amplitude tolerances of BASELINE.json's north_star, relative to max|a|
TOL = {np.float64: 1e-12, np.float32: 1e-5}


def rel_err(actual, expected):
    expected = np.asarray(expected)
    scale = max(np.abs(expected).max(), 1e-300)
    return np.abs(np.asarray(actual) - expected).max() / scale


def make_sim(runtime, dtype, prep):
    return qgate_b200.simulator.with_runtime(runtime, dtype=dtype, circuit_prep=prep)


def run_circuit(runtime, name, dtype, prep):
    sim = make_sim(runtime, dtype, prep)
    q, ops = CIRCUITS[name]()
    sim.run(ops)
    sim.qubits.set_ordering(q)
    return sim, q


def midcircuit_ops():
    q = S.new_qregs(5)
    refs = S.new_references(3)
    ops = [S.H(qr) for qr in q] + [S.ctrl(q[0]).X(q[1]), S.ctrl(q[1]).Ry(0.4)(q[2]),
           S.measure(refs[0], q[1]), S.reset(q[1]), S.H(q[1]),
           S.ctrl(q[1], q[2]).X(q[3]), S.measure(refs[1], q[3]),
           S.if_(refs[1], 1, [S.X(q[4]), S.ctrl(q[4]).H(q[0])]),
           S.if_([refs[0], refs[1]], lambda a, b: a == b, S.Z(q[2])),
           S.ctrl(q[2]).U3(0.3, 0.2, 0.1)(q[4]), S.measure(refs[2], q[0])]
    return q, refs, ops


GATE_SPECS = [('I', ()), ('H', ()), ('S', ()), ('T', ()), ('X', ()), ('Y', ()), ('Z', ()),
              ('SH', ()), ('Rx', (0.3,)), ('Ry', (-1.1,)), ('Rz', (2.2,)), ('U1', (0.7,)),
              ('U2', (0.4, -0.9)), ('U3', (1.3, 0.2, -2.5)), ('Expii', (0.6,)),
              ('Expiz', (-0.8,))]
SCRIPT_TO_GATE_ID = {'I': 'ID', 'Rx': 'RX', 'Ry': 'RY', 'Rz': 'RZ', 'U3': 'U', 'Expii': 'ExpiI',
                     'Expiz': 'ExpiZ'}
