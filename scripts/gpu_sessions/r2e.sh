#!/bin/bash
# round-2 GPU session e: whole -m gpu suite after the lane map / native multi-qubit ops / native sample,
# default bench line (short), Simulator.sample timing native vs loop.
mkdir -p gpurun_out
( time timeout 1800 python -m pytest tests -m gpu -q --durations=8 ) > gpurun_out/r2e_pytest.log 2>&1; tail -30 gpurun_out/r2e_pytest.log
timeout 600 python bench.py --steps 3 --warmup 3 --depth 60 --no-cpu-baseline > gpurun_out/r2e_bench_d60.json 2> gpurun_out/r2e_bench_d60.err; python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/r2e_bench_d60.json').read().strip().splitlines()[-1])
    r = d['roofline']
    print('f64 upd/s %.3e frac %.3f pipe %.3f passes %.0f | e2e %.3e (%s) | f32 %.3e frac %.3f | qft %s' % (d['value'], r['frac'], r['pipe']['frac'], r['launches_per_step'], d['e2e']['value'], d['e2e']['front_end'], d['f32']['value'], d['f32']['roofline']['frac'], [(q['qubits'], round(q.get('ms', -1), 1), q['ok']) for q in d['qft']]))
except Exception as e:
    print('bench failed', e, open('gpurun_out/r2e_bench_d60.err').read()[-600:])
PY
python - <<'PY'
import time, numpy as np, sys
sys.path.insert(0, '.')
import qgate_b200, qgate_b200.script as S
from qgate_b200 import circuits, cudaruntime
from qgate_b200.simulator import with_runtime
q, ops = circuits.random_u3_cx(S, 16, 10, seed=3)
refs = S.new_references(16)
ops = ops + [S.measure(r, x) for r, x in zip(refs, q)]
for native in (True, False):
    sim = with_runtime(cudaruntime, dtype=np.float64, circuit_prep='one_static', native_sample=native)
    np.random.seed(1)
    shots = 2000 if native else 200
    t0 = time.perf_counter()
    obs = sim.sample(ops, refs, shots).intarray
    dt = time.perf_counter() - t0
    print('Simulator.sample 16 qubits x 10 layers, measure all: native=%s  %d shots in %.3f s = %.2f ms/shot' % (native, shots, dt, 1e3 * dt / shots))
    sim.terminate()
PY
