#!/bin/bash
# round-2 GPU session c: the whole -m gpu suite (new: exact mode, reference-compatible pool, survey-size
# oracle comparisons, the reference front end on the CUDA library), bench line with sub-records,
# default sweeps for both precisions, ncu captures of the sheared kernel.
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q --durations=15 ) > gpurun_out/r2c_pytest.log 2>&1; tail -40 gpurun_out/r2c_pytest.log
( QGB_MEAS_TOL_MULT=1 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "measure_all or midcircuit" ) > gpurun_out/r2c_pytest_meas_tol1.log 2>&1; tail -15 gpurun_out/r2c_pytest_meas_tol1.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r2c_bench_default.json 2> gpurun_out/r2c_bench_default.err; tail -c 6000 gpurun_out/r2c_bench_default.json; tail -5 gpurun_out/r2c_bench_default.err
run() {
  tag=$1; shift
  timeout 300 python bench.py --steps 3 --warmup 3 --depth 60 --no-e2e --no-cpu-baseline --no-extras "$@" > gpurun_out/r2c_$tag.json 2> gpurun_out/r2c_$tag.err
  python - "$tag" gpurun_out/r2c_$tag.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    r = d['roofline']
    print('%-28s ms/pass %.3f frac %.3f pipe %.3f upd/s %.3e passes %.0f sm %s W %s %s' % (sys.argv[1], r['avg_launch_ms'], r['frac'], r['pipe']['frac'], d['value'], r['launches_per_step'], d['clocks']['sm_mhz'], d['clocks'].get('power_w_max'), d['clocks']['reasons']))
except Exception as e:
    print(sys.argv[1], 'failed', e, open(sys.argv[2].replace('.json', '.err')).read()[-400:])
PY
}
for T in 10 11; do for c in 21 24 27; do run f64_T${T}_c$c --option tile_lanes_fp64=$T --option max_cost=$c; done; done
run f64_T11_c24_b3 --option tile_lanes_fp64=11 --option max_cost=24 --option tma_buffers=3
for T in 11 12; do for c in 21 24 27 33; do run f32_T${T}_c$c --dtype f32 --option tile_lanes_fp32=$T --option max_cost=$c; done; done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2c_launches.csv python bench.py --steps 2 --warmup 1 --depth 20 --no-e2e --no-cpu-baseline --no-extras > gpurun_out/r2c_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tma_pass -s 6 -c 2 -o gpurun_out/r2c_tma_f64_30q -f python bench.py --steps 1 --warmup 1 --depth 10 --no-e2e --no-cpu-baseline --no-extras > gpurun_out/r2c_ncu_f64.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tma_pass -s 6 -c 2 -o gpurun_out/r2c_tma_f32_30q -f python bench.py --dtype f32 --steps 1 --warmup 1 --depth 10 --no-e2e --no-cpu-baseline --no-extras > gpurun_out/r2c_ncu_f32.log 2>&1
tail -2 gpurun_out/r2c_ncu_f64.log
