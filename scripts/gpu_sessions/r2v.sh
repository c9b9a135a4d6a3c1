#!/bin/bash
# round-2 GPU session v: lazy |0...0> (the first fused pass starts from the basis state without reading the
# array).  Whole -m gpu suite, QFT timing with / without, short bench.
mkdir -p gpurun_out
( time timeout 1800 python -m pytest tests -m gpu -q -x --durations=4 ) > gpurun_out/r2v_pytest.log 2>&1; tail -10 gpurun_out/r2v_pytest.log
timeout 300 python tools/qft_breakdown.py 30 3 2>&1 | tail -1
timeout 300 python tools/qft_breakdown.py 30 3 lazy_reset=0 2>&1 | tail -1
timeout 300 python tools/qft_breakdown.py 32 3 2>&1 | tail -1
timeout 300 python tools/qft_breakdown.py 32 3 lazy_reset=0 2>&1 | tail -1
timeout 600 python bench.py --steps 3 --warmup 3 --depth 60 --no-cpu-baseline > gpurun_out/r2v_bench_d60.json 2> gpurun_out/r2v_bench_d60.err; python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/r2v_bench_d60.json').read().strip().splitlines()[-1])
    r = d['roofline']
    print('f64 upd/s %.3e frac %.3f pipe %.3f passes %.0f | e2e %.3e (%s) | f32 %.3e frac %.3f | qft %s' % (d['value'], r['frac'], r['pipe']['frac'], r['launches_per_step'], d['e2e']['value'], d['e2e']['front_end'], d['f32']['value'], d['f32']['roofline']['frac'], [(q['qubits'], round(q.get('ms', -1), 1), round(q.get('first_run_ms', -1), 1), q['ok']) for q in d['qft']]))
except Exception as e:
    print('bench failed', e, open('gpurun_out/r2v_bench_d60.err').read()[-600:])
PY
