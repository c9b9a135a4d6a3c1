#!/bin/bash
# round-2 GPU session t: 9-lane tiles (one-warp CTAs, 8 KiB tiles) against the 10-lane default, complex128;
# 10-lane tiles for complex64.
mkdir -p gpurun_out
run() {
  tag=$1; shift
  timeout 300 python bench.py --steps 3 --warmup 3 --depth 60 --no-e2e --no-cpu-baseline --no-extras "$@" > gpurun_out/r2t_$tag.json 2> gpurun_out/r2t_$tag.err
  python - "$tag" gpurun_out/r2t_$tag.json <<'PY'
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
run f64_T9_c24 --option tile_lanes_fp64=9
run f64_T9_c21 --option tile_lanes_fp64=9 --option max_cost=21
run f64_T9_c18 --option tile_lanes_fp64=9 --option max_cost=18
run f64_T9_c21_b3 --option tile_lanes_fp64=9 --option max_cost=21 --option tma_buffers=3
run f64_T9_c21_wl --option tile_lanes_fp64=9 --option max_cost=21 --option warp_local=1
run f32_default --dtype f32
run f32_T10_c21 --dtype f32 --option tile_lanes_fp32=10
run f32_T10_c18 --dtype f32 --option tile_lanes_fp32=10 --option max_cost=18
