#!/bin/bash
# round-2 GPU session i: phase fans with tile factors from device tables (fan_tile_table_kernel); fan_cost / max_cost sweep on
# QFT-30 and QFT-33; ncu launch list and one full capture of a QFT pass.
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "fan or qft or phase_estimation" ) > gpurun_out/r2i_pytest_fans.log 2>&1; tail -3 gpurun_out/r2i_pytest_fans.log
run() {
  n=$1; tag=$2; shift; shift
  timeout 600 python run_configs.py qft --repeat 2 --qubits $n "$@" > gpurun_out/r2i_qft${n}_$tag.json 2> gpurun_out/r2i_qft${n}_$tag.err
  python - $n $tag <<'PY'
import json, sys
n, tag = sys.argv[1:3]
try:
    d = json.loads(open('gpurun_out/r2i_qft%s_%s.json' % (n, tag)).read().strip().splitlines()[-1])
    st = d.get('engine_stats', {})
    print('QFT-%s %-18s: run %.1f ms, passes %s, fan ops %s, rel err %.2e, ok %s' % (n, tag, 1e3 * d['run_s'], st.get('tile_passes'), st.get('fan_ops'), d['amplitude_rel_err_vs_closed_form'], d['ok']))
except Exception as e:
    print('qft', n, tag, 'failed', e, open('gpurun_out/r2i_qft%s_%s.err' % (n, tag)).read()[-800:])
PY
}
for n in 30 33; do
  run $n default
  run $n fc1 --option fan_cost=1
  run $n fc1_c30 --option fan_cost=1 --option max_cost=30
  run $n fc2_c30 --option fan_cost=2 --option max_cost=30
  run $n fc2_c36 --option fan_cost=2 --option max_cost=36
  run $n fc1_c24_T11 --option fan_cost=1 --option tile_lanes_fp64=11
done
run 30 f32 --dtype f32
run 30 f32_fc1 --dtype f32 --option fan_cost=1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 100 --csv --log-file gpurun_out/r2i_qft30_launches.csv python run_configs.py qft --qubits 30 > gpurun_out/r2i_qft30_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tma_pass -s 8 -c 2 -o gpurun_out/r2i_qft30_pass -f python run_configs.py qft --qubits 30 > gpurun_out/r2i_ncu_qft.log 2>&1
tail -2 gpurun_out/r2i_ncu_qft.log
grep tma_pass gpurun_out/r2i_qft30_launches.csv | awk -F'","' '{print $NF}' | tr -d '"' | tr '\n' ' '
