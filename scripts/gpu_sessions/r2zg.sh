#!/bin/bash
# round-2 GPU session zg (last GPU minutes): the driver's bench command with the final code
mkdir -p gpurun_out
timeout 170 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2zg_bench_driver_cmd.json 2> gpurun_out/r2zg_bench_driver_cmd.err
python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/r2zg_bench_driver_cmd.json').read().strip().splitlines()[-1])
    r = d['roofline']
    print('f64 upd/s %.3e ms/step %.1f frac %.3f pipe %.3f | e2e %.3e | f32 %.3e frac %.3f | qft %s' % (d['value'], d['ms_per_step'], r['frac'], r['pipe']['frac'], d['e2e']['value'], d['f32']['value'], d['f32']['roofline']['frac'], [(q['qubits'], round(q.get('ms', -1), 1), q['ok']) for q in d['qft']]))
except Exception as e:
    print('bench failed', e, open('gpurun_out/r2zg_bench_driver_cmd.err').read()[-600:])
PY
