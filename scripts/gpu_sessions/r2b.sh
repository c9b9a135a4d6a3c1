#!/bin/bash
# round-2 GPU session b: sheared ops (OP_SHEAR) — parity suite, then cost / hint sweeps and the
# energy attribution of a pass (debug_pass_mode 1: stages without ops, 2: TMA in/out only).
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2b_pytest.log 2>&1; tail -5 gpurun_out/r2b_pytest.log
run() {
  tag=$1; shift
  timeout 300 python bench.py --steps 3 --warmup 3 --depth 60 --no-e2e --no-cpu-baseline "$@" > gpurun_out/r2b_$tag.json 2> gpurun_out/r2b_$tag.err
  python - "$tag" gpurun_out/r2b_$tag.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    r = d['roofline']
    print('%-28s ms/pass %.3f frac %.3f upd/s %.3e passes %.0f sm %s W %s %s p0 %.6f' % (sys.argv[1], r['avg_launch_ms'], r['frac'], d['value'], r['launches_per_step'], d['clocks']['sm_mhz'], d['clocks'].get('power_w_max'), d['clocks']['reasons'], d['p0_check']))
except Exception as e:
    print(sys.argv[1], 'failed', e, open(sys.argv[2].replace('.json', '.err')).read()[-400:])
PY
}
run noshear_c24 --option shear=0 --option max_cost=24
for c in 15 18 21 24 27 30 36; do run shear_c$c --option max_cost=$c; done
for h in 1 2 3; do run shear_c21_hint$h --option max_cost=21 --option l2_hint=$h; done
run dbg1_c21 --option max_cost=21 --option debug_pass_mode=1
run dbg2_c21 --option max_cost=21 --option debug_pass_mode=2
run shear_c24_T11 --option max_cost=24 --option tile_lanes_fp64=11
run shear_c30_T11 --option max_cost=30 --option tile_lanes_fp64=11
run shear_c21_k3 --option max_cost=21 --option reg_bits_fp64=3
run f32_noshear --dtype f32
run f32_shear --dtype f32 --option shear_fp32=1
