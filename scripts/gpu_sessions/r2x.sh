#!/bin/bash
# round-2 GPU session x: queue streaming (passes launched while gates are still being submitted).  Whole
# -m gpu suite, then the driver's bench command.
mkdir -p gpurun_out
( time timeout 1800 python -m pytest tests -m gpu -q -x --durations=4 ) > gpurun_out/r2x_pytest.log 2>&1; tail -8 gpurun_out/r2x_pytest.log
( time timeout 1200 python bench.py --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/r2x_bench_driver_cmd.json 2> gpurun_out/r2x_bench_driver_cmd.err; tail -3 gpurun_out/r2x_bench_driver_cmd.err
python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/r2x_bench_driver_cmd.json').read().strip().splitlines()[-1])
    r = d['roofline']
    print('f64 upd/s %.3e ms/step %.1f frac %.3f pipe %.3f passes %.0f clocks %s | e2e %.3e (%s) split %s | f32 %.3e frac %.3f | qft %s' % (d['value'], d['ms_per_step'], r['frac'], r['pipe']['frac'], r['launches_per_step'], d['clocks'], d['e2e']['value'], d['e2e']['front_end'], d['e2e']['split_ms'], d['f32']['value'], d['f32']['roofline']['frac'], [(q['qubits'], round(q.get('ms', -1), 1), q['ok']) for q in d['qft']]))
except Exception as e:
    print('bench failed', e, open('gpurun_out/r2x_bench_driver_cmd.err').read()[-600:])
PY
