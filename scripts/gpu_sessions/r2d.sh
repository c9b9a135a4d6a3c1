#!/bin/bash
# round-2 GPU session d: ranks sharing one GPU (gloo control plane + CUDA-IPC exchange, push and in place),
# the fixed survey-size tests, Grover-30 pool build timing, low-lane experiments.
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests/test_dist_shared_gpu.py -m gpu -q -x ) > gpurun_out/r2d_pytest_shared.log 2>&1; tail -25 gpurun_out/r2d_pytest_shared.log
( time timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "grover_22 or phase_estimation_20 or exact or compatible or measure_all or midcircuit or fused_equals" ) > gpurun_out/r2d_pytest_parity.log 2>&1; tail -15 gpurun_out/r2d_pytest_parity.log
for opt in "pool_from_amplitudes=1" "pool_from_amplitudes=0"; do
  timeout 300 python run_configs.py grover --qubits 30 --option $opt > gpurun_out/r2d_grover30_$opt.json 2> gpurun_out/r2d_grover30_$opt.err; python -c "
import json,sys
d=json.loads(open('gpurun_out/r2d_grover30_$opt.json').read().strip().splitlines()[-1])
print('$opt', {k:d[k] for k in ('ok','run_s','pool_build_s','sample_s','measure_all_s','p_marked','hits_marked')})" || tail -3 gpurun_out/r2d_grover30_$opt.err
done
run() {
  tag=$1; shift
  timeout 300 python bench.py --steps 3 --warmup 3 --depth 60 --no-e2e --no-cpu-baseline --no-extras "$@" > gpurun_out/r2d_$tag.json 2> gpurun_out/r2d_$tag.err
  python - "$tag" gpurun_out/r2d_$tag.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    r = d['roofline']
    print('%-28s ms/pass %.3f frac %.3f pipe %.3f upd/s %.3e passes %.0f sm %s W %s %s' % (sys.argv[1], r['avg_launch_ms'], r['frac'], r['pipe']['frac'], d['value'], r['launches_per_step'], d['clocks']['sm_mhz'], d['clocks'].get('power_w_max'), d['clocks']['reasons']))
except Exception as e:
    print(sys.argv[1], 'failed', e, open(sys.argv[2].replace('.json', '.err')).read()[-400:])
PY
}
run f64_default
run f64_L4 --option low_lanes_fp64=4
run f64_L4_c27 --option low_lanes_fp64=4 --option max_cost=27
run f64_L4_c30 --option low_lanes_fp64=4 --option max_cost=30
run f64_L3_c27 --option low_lanes_fp64=3 --option max_cost=27
run f64_L6 --option low_lanes_fp64=6
run f32_default --dtype f32
run f32_L5 --dtype f32 --option low_lanes_fp32=5
run f32_L5_c24 --dtype f32 --option low_lanes_fp32=5 --option max_cost=24
