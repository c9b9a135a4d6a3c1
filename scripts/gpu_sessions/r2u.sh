#!/bin/bash
# round-2 GPU session u: what the driver runs at round end, with the final code: the whole -m gpu suite,
# smoke(), the bench command (20 steps) and the reference arm.
mkdir -p gpurun_out
( time timeout 1800 python -m pytest tests -m gpu -q --durations=8 ) > gpurun_out/r2u_pytest.log 2>&1; tail -16 gpurun_out/r2u_pytest.log
( time timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" ) 2>&1 | tail -5
( time timeout 1200 python bench.py --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/r2u_bench_driver_cmd.json 2> gpurun_out/r2u_bench_driver_cmd.err; tail -3 gpurun_out/r2u_bench_driver_cmd.err
python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/r2u_bench_driver_cmd.json').read().strip().splitlines()[-1])
    r = d['roofline']
    print('f64 upd/s %.3e ms/step %.1f frac %.3f pipe %.3f passes %.0f clocks %s | e2e %.3e (%s) | f32 %.3e frac %.3f | qft %s | cpu %s' % (d['value'], d['ms_per_step'], r['frac'], r['pipe']['frac'], r['launches_per_step'], d['clocks'], d['e2e']['value'], d['e2e']['front_end'], d['f32']['value'], d['f32']['roofline']['frac'], [(q['qubits'], round(q.get('ms', -1), 1), q['ok']) for q in d['qft']], d['cpu_baseline']))
except Exception as e:
    print('bench failed', e, open('gpurun_out/r2u_bench_driver_cmd.err').read()[-600:])
PY
( time timeout 900 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 ) > gpurun_out/r2u_bench_reference_arm.json 2> gpurun_out/r2u_bench_reference_arm.err; tail -c 900 gpurun_out/r2u_bench_reference_arm.json
