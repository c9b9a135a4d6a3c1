#!/usr/bin/env python
"""Generates tests/golden/golden_v1.npz from the REAL reference (runs only in the build
container, where /root/reference exists; the fixtures travel, the reference does not).

Recipe: copy /root/reference to a scratch dir, build its glue.so + cpuext.so
with its own Makefile targets, stub qgate.openqasm (PLY is absent), import qgate, and run the
circuits of qgate_b200/circuits.py through qgate.simulator.py (complex128 oracle) and
qgate.simulator.cpu (float64 and float32).  Nothing of the reference is copied into the repo.

    python tests/golden/make_golden.py            # ~1 min
    python tests/golden/make_golden.py --qft20    # adds the 20-qubit py-runtime run (~7 min)
"""
import argparse
import os
import subprocess
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
SCRATCH = '/tmp/qgate_ref'


def import_reference():
    if not os.path.exists(os.path.join(SCRATCH, 'qgate', 'simulator', 'cpuext.so')):
        subprocess.check_call('rm -rf {0} && cp -r /root/reference {0} && chmod -R u+w {0}'
                              .format(SCRATCH), shell=True)
        src = os.path.join(SCRATCH, 'qgate', 'simulator', 'src')
        subprocess.check_call('python3 incpathgen.py > incpath && make ../glue.so ../cpuext.so -j8',
                              shell=True, cwd=src, stdout=subprocess.DEVNULL,
                              stderr=subprocess.DEVNULL)
    sys.path.insert(0, SCRATCH)
    sys.modules['qgate.openqasm'] = types.ModuleType('qgate.openqasm')
    import qgate
    return qgate


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--qft20', action='store_true')
    args = ap.parse_args()

    os.environ.setdefault('QGATE_NUM_WORKERS', '4')   # pin: calc_probability depends on it (~1e-14)
    qgate = import_reference()
    import qgate.script as S
    sys.path.insert(0, REPO)
    from qgate_b200 import circuits

    out = dict()

    def sims(prep):
        return [('py', qgate.simulator.py(circuit_prep=prep)),
                ('cpu64', qgate.simulator.cpu(circuit_prep=prep, dtype=np.float64)),
                ('cpu32', qgate.simulator.cpu(circuit_prep=prep, dtype=np.float32))]

    # ---- A. gate matrices: columns read back from 1-qubit states on the cpu runtime ------
    gate_specs = [('I', ()), ('H', ()), ('S', ()), ('T', ()), ('X', ()), ('Y', ()), ('Z', ()),
                  ('SH', ()), ('Rx', (0.3,)), ('Ry', (-1.1,)), ('Rz', (2.2,)), ('U1', (0.7,)),
                  ('U2', (0.4, -0.9)), ('U3', (1.3, 0.2, -2.5)), ('Expii', (0.6,)),
                  ('Expiz', (-0.8,))]
    for name, params in gate_specs:
        for adj in (False, True):
            cols = []
            for basis in (0, 1):
                q = S.new_qreg()
                factory = getattr(S, name)
                if params:
                    factory = factory(*params)
                if adj:
                    factory = factory.Adj
                ops = ([S.X(q)] if basis else [S.I(q)]) + [factory(q)]
                sim = qgate.simulator.cpu(dtype=np.float64)
                sim.run(ops)
                cols.append(sim.qubits.states[:])
            out['matrix/{}/{}'.format(name, 'adj' if adj else 'fwd')] = np.stack(cols, axis=1)
            out['matrix_args/{}'.format(name)] = np.array(params, np.float64)

    # ---- B/C. circuits: amplitudes, probabilities, per-qreg P(0) ---------------------------
    cases = {
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
    for name, make in cases.items():
        for prep in ('dynamic', 'one_static'):
            for rt, sim in sims(prep):
                q, ops = make()
                sim.run(ops)
                sim.qubits.set_ordering(q)
                key = '{}/{}/{}'.format(name, prep, rt)
                out[key + '/states'] = sim.qubits.states[:]
                out[key + '/prob'] = sim.qubits.prob[:]
                out[key + '/p0'] = np.array([sim.qubits.calc_probability(qr) for qr in q])

    # phase estimation (target is a hidden lane of the pool)
    for rt, sim in sims('dynamic'):
        bits, target, ops = circuits.phase_estimation(S, 8, 0.1)
        sim.run(ops)
        sim.qubits.set_ordering(bits + [target])
        out['pe8/{}/states'.format(rt)] = sim.qubits.states[:]
        rnd = np.random.RandomState(3).random_sample(4096)
        pool = sim.qubits.create_sampling_pool(bits)
        out['pe8/{}/samples'.format(rt)] = pool.sample(4096, rnd).intarray

    # ---- D. measurement: outcomes + collapsed states under a fixed global seed -------------
    for name in ('rand6x10', 'zoo7', 'grover8', 'ghz11'):
        for prep in ('dynamic', 'one_static'):
            for rt, sim in sims(prep):
                q, ops = cases[name]()
                refs = S.new_references(len(q))
                ops = ops + [S.measure(r, qr) for r, qr in zip(refs, q)]
                np.random.seed(11)
                sim.run(ops)
                sim.qubits.set_ordering(q)
                key = 'measure/{}/{}/{}'.format(name, prep, rt)
                out[key + '/bits'] = np.array(sim.values.get(refs), np.int64)
                out[key + '/states'] = sim.qubits.states[:]
    # measure in the middle, keep computing, reset, if-clause
    for prep in ('dynamic', 'one_static'):
        for rt, sim in sims(prep):
            q = S.new_qregs(5)
            refs = S.new_references(3)
            ops = [S.H(qr) for qr in q] + [S.ctrl(q[0]).X(q[1]), S.ctrl(q[1]).Ry(0.4)(q[2]),
                   S.measure(refs[0], q[1]), S.reset(q[1]), S.H(q[1]),
                   S.ctrl(q[1], q[2]).X(q[3]), S.measure(refs[1], q[3]),
                   S.if_(refs[1], 1, [S.X(q[4]), S.ctrl(q[4]).H(q[0])]),
                   S.if_([refs[0], refs[1]], lambda a, b: a == b, S.Z(q[2])),
                   S.ctrl(q[2]).U3(0.3, 0.2, 0.1)(q[4]), S.measure(refs[2], q[0])]
            np.random.seed(5)
            sim.run(ops)
            sim.qubits.set_ordering(q)
            key = 'measure/midcircuit/{}/{}'.format(prep, rt)
            out[key + '/bits'] = np.array(sim.values.get(refs), np.int64)
            out[key + '/states'] = sim.qubits.states[:]

    # ---- E. sampling pool: fixed randnum, hidden lanes, empty lanes, reordering --------------
    rnd = np.random.RandomState(7).random_sample(20000)
    out['sampling/randnum_seed'] = np.array([7, 20000])
    for name in ('rand10x20', 'zoo9', 'grover8'):
        for prep in ('dynamic', 'one_static'):
            for rt, sim in sims(prep):
                q, ops = cases[name]()
                empty = S.new_qregs(2)
                sim.run(ops)
                key = 'sampling/{}/{}/{}'.format(name, prep, rt)
                out[key + '/full'] = sim.qubits.create_sampling_pool(q).sample(20000, rnd).intarray
                sub = q[1::2]
                out[key + '/hidden'] = sim.qubits.create_sampling_pool(sub).sample(20000, rnd).intarray
                mixed = [q[3], empty[0], q[0], q[5], empty[1], q[1]]
                out[key + '/empty'] = sim.qubits.create_sampling_pool(mixed).sample(20000, rnd).intarray
                rev = list(reversed(q))
                out[key + '/reversed'] = sim.qubits.create_sampling_pool(rev).sample(20000, rnd).intarray

                class Probe:
                    def __init__(self, prob, empty_lanes, qreg_ordering):
                        self.prob = prob
                out[key + '/prob_hidden'] = sim.qubits.create_sampling_pool(sub, Probe).prob

    # ---- F. get_states slicing ------------------------------------------------------------------
    slice_keys = [(None, None, None), (3, 200, 7), (None, None, -1), (500, 20, -3), (-5, None, 1),
                  (10, 11, 1), (0, 1024, 300), (1023, None, -1023), (-1, -1025, -1)]
    out['slices/keys'] = np.array([[-(1 << 40) if v is None else v for v in k] for k in slice_keys])
    for rt, sim in sims('dynamic'):
        q, ops = cases['rand10x20']()
        extra = S.new_qregs(1)
        sim.run(ops)
        sim.qubits.set_ordering(q)
        for idx, k in enumerate(slice_keys):
            out['slices/{}/{}/states'.format(rt, idx)] = sim.qubits.states[slice(*k)]
            out['slices/{}/{}/prob'.format(rt, idx)] = sim.qubits.prob[slice(*k)]
        # ordering with an unallocated (empty) qreg in the middle
        sim.qubits.set_ordering(q[:4] + extra + q[4:])
        out['slices/{}/empty_lane/states'.format(rt)] = sim.qubits.states[::5]

    # ---- config 0: QFT-20 complex128 -------------------------------------------------------------
    q, ops = circuits.qft(S, 20)
    sim = qgate.simulator.cpu(dtype=np.float64, circuit_prep='one_static')
    sim.run(ops)
    sim.qubits.set_ordering(q)
    out['qft20/cpu64/states_stride4099'] = sim.qubits.states[::4099]
    out['qft20/cpu64/states_head'] = sim.qubits.states[:512]
    out['qft20/cpu64/p0'] = np.array([sim.qubits.calc_probability(qr) for qr in q])
    full = sim.qubits.states[:]
    out['qft20/cpu64/checksum'] = np.array([np.vdot(full, full).real, full.sum().real,
                                            full.sum().imag, np.abs(full).max()])
    if args.qft20:
        q, ops = circuits.qft(S, 20)
        sim = qgate.simulator.py(circuit_prep='one_static')
        sim.run(ops)
        sim.qubits.set_ordering(q)
        out['qft20/py/states_stride4099'] = sim.qubits.states[::4099]
        out['qft20/py/states_head'] = sim.qubits.states[:512]
    else:
        prev = os.path.join(HERE, 'golden_v1.npz')
        if os.path.exists(prev):
            with np.load(prev) as old:
                for k in old.files:
                    if k.startswith('qft20/py/'):
                        out[k] = old[k]

    np.savez_compressed(os.path.join(HERE, 'golden_v1.npz'), **out)
    print('wrote {} arrays'.format(len(out)))


if __name__ == '__main__':
    main()
