"""CPU suite: the parts of bench.py's contract that need no GPU — the reference arm prints one JSON
line with the agreed keys (and only rank 0 does under a multi-rank launch), and the workload
generator / update count agree with qgate_b200.circuits."""
import json
import os
import subprocess
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    out = subprocess.run([sys.executable, os.path.join(REPO, 'bench.py')] + args, stdout=subprocess.PIPE,
                         stderr=subprocess.PIPE, text=True, env=e, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    return out.stdout


def test_reference_arm_line(ref_runtime):
    out = _run(['--impl', 'reference', '--qubits', '14', '--steps', '2', '--warmup', '1'])
    lines = [l for l in out.splitlines() if l.startswith('{')]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['metric'] == 'gate-amp updates/s' and d['unit'] == 'updates/s'
    assert d['higher_is_better'] is True and d['n_gpus'] == 1 and d['steps'] == 2 and d['warmup'] == 1
    assert d['value'] > 0 and d['e2e']['value'] == d['value']
    assert d['e2e']['h2d_bytes_per_step'] == 0 and d['e2e']['d2h_bytes_per_step'] == 0
    assert d['cpu_baseline']['kind'] == 'reference' and d['cpu_baseline']['cores'] >= 1
    assert d['cpu_baseline']['value'] == d['value'] and 'workload' in d['config']
    assert d['gpu_launches'] == 0 and d['vs_baseline'] is None


def test_reference_arm_other_ranks_stay_silent(ref_runtime):
    out = _run(['--impl', 'reference', '--qubits', '12', '--gpus', '2', '--steps', '1', '--warmup', '1'],
               env={'RANK': '1', 'LOCAL_RANK': '1', 'WORLD_SIZE': '2'})
    assert out.strip() == ''


def test_workload_matches_circuit_generator():
    sys.path.insert(0, REPO)
    import bench
    import qgate_b200.script as S
    from qgate_b200 import circuits, model
    n, depth = 7, 5
    gates = bench.random_circuit_gates(n, depth, 1234)
    q, ops = circuits.random_u3_cx(S, n, depth, seed=1234)
    assert len(gates) == len(ops) == circuits.random_circuit_gate_count(n, depth)
    lane = {qr.id: i for i, qr in enumerate(q)}
    for (angles, ctrl, target), op in zip(gates, ops):
        assert lane[op.qreg.id] == target
        if ctrl < 0:
            assert op.ctrllist is None and np.allclose(op.gate_type.args, angles, rtol=0, atol=0)
        else:
            assert [lane[c.id] for c in op.ctrllist] == [ctrl]
    assert bench.updates_of(gates, n) == sum((1 << n) if g[1] < 0 else (1 << (n - 1)) for g in gates)
