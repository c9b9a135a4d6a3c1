#!/bin/bash
# round-2 GPU session m: fan bodies with independent per-register factors, fan code only in the FANS kernel variant.  Parity tests, QFT timings, launch list,
# bench (short) to check the main path did not move.
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "fan or qft or phase_estimation or fused_equals" ) > gpurun_out/r2m_pytest_fans.log 2>&1; tail -3 gpurun_out/r2m_pytest_fans.log
timeout 300 python tools/qft_breakdown.py 30 3 2>&1 | tail -2
timeout 300 python tools/qft_breakdown.py 30 3 fan_cost=1 2>&1 | tail -1
timeout 300 python tools/qft_breakdown.py 30 3 fan_cost=1 max_cost=30 2>&1 | tail -1
timeout 300 python tools/qft_breakdown.py 32 3 2>&1 | tail -1
timeout 300 python tools/qft_breakdown.py 32 3 fan_cost=1 2>&1 | tail -1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 100 --csv --log-file gpurun_out/r2m_qft30_launches.csv python run_configs.py qft --qubits 30 > gpurun_out/r2m_qft30_launches.log 2>&1
echo "QFT-30 per-pass ns:"; grep tma_pass gpurun_out/r2m_qft30_launches.csv | awk -F'","' '{print $NF}' | tr -d '"' | tr '\n' ' '; echo
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tma_pass -s 1 -c 1 -o gpurun_out/r2m_qft30_pass -f python run_configs.py qft --qubits 30 > gpurun_out/r2m_ncu_qft30.log 2>&1
timeout 600 python bench.py --steps 3 --warmup 3 --depth 60 --no-cpu-baseline > gpurun_out/r2m_bench_d60.json 2> gpurun_out/r2m_bench_d60.err; python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/r2m_bench_d60.json').read().strip().splitlines()[-1])
    r = d['roofline']
    print('f64 upd/s %.3e frac %.3f pipe %.3f passes %.0f | e2e %.3e (%s) | f32 %.3e frac %.3f | qft %s' % (d['value'], r['frac'], r['pipe']['frac'], r['launches_per_step'], d['e2e']['value'], d['e2e']['front_end'], d['f32']['value'], d['f32']['roofline']['frac'], [(q['qubits'], round(q.get('ms', -1), 1), q['ok']) for q in d['qft']]))
except Exception as e:
    print('bench failed', e, open('gpurun_out/r2m_bench_d60.err').read()[-600:])
PY
