#!/bin/bash
# round-2 GPU session g: phase fans (OP_FAN).  Fan parity tests + every QFT / phase-estimation test,
# QFT-30 / QFT-33 with and without fans, then the driver's own bench command (20 steps, depth 200).
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --durations=6 -k "fan or qft or phase_estimation or large_state or against_reference_cpu_runtime or fused_equals" ) > gpurun_out/r2g_pytest_fans.log 2>&1; tail -14 gpurun_out/r2g_pytest_fans.log
for fans in 1 0; do
  for n in 30 33; do
    timeout 600 python run_configs.py qft --qubits $n --option fans=$fans > gpurun_out/r2g_qft${n}_fans${fans}.json 2> gpurun_out/r2g_qft${n}_fans${fans}.err
    python - $n $fans <<'PY'
import json, sys
n, fans = sys.argv[1:3]
try:
    d = json.loads(open('gpurun_out/r2g_qft%s_fans%s.json' % (n, fans)).read().strip().splitlines()[-1])
    st = d.get('engine_stats', {})
    print('QFT-%s fans=%s: run %.1f ms, passes %s, fan ops %s, rel err %.2e, ok %s' % (n, fans, 1e3 * d['run_s'], st.get('tile_passes'), st.get('fan_ops'), d['amplitude_rel_err_vs_closed_form'], d['ok']))
except Exception as e:
    print('qft', n, fans, 'failed', e, open('gpurun_out/r2g_qft%s_fans%s.err' % (n, fans)).read()[-800:])
PY
  done
done
( time timeout 1200 python bench.py --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/r2g_bench_driver_cmd.json 2> gpurun_out/r2g_bench_driver_cmd.err; tail -3 gpurun_out/r2g_bench_driver_cmd.err
python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/r2g_bench_driver_cmd.json').read().strip().splitlines()[-1])
    r = d['roofline']
    print('f64 upd/s %.3e ms/step %.1f frac %.3f pipe %.3f passes %.0f clocks %s | e2e %.3e (%s) | f32 %.3e frac %.3f | qft %s | cpu %s' % (d['value'], d['ms_per_step'], r['frac'], r['pipe']['frac'], r['launches_per_step'], d['clocks'], d['e2e']['value'], d['e2e']['front_end'], d['f32']['value'], d['f32']['roofline']['frac'], [(q['qubits'], round(q.get('ms', -1), 1), q['ok']) for q in d['qft']], d['cpu_baseline']))
except Exception as e:
    print('bench failed', e, open('gpurun_out/r2g_bench_driver_cmd.err').read()[-600:])
PY
