#!/bin/bash
# round-1 GPU session ZB: (contiguous run length) x (pass cost cap) sweep, both precisions
mkdir -p gpurun_out
Q="--steps 1 --warmup 1 --no-e2e --no-cpu-baseline --depth 60 --option tma_ws=0"
show() { python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()[:300]); continue
    r=d['roofline']; print('value %.3e passes %d gates/pass %.1f avg_ms %.2f GB/s %.0f frac %.3f clk %s'%(d['value'],r['launches_per_step'],r['gates_per_launch'],r['avg_launch_ms'],r['achieved'],r['frac'],d['clocks']['sm_mhz']))
"; }
for L in 4 5 6; do for C in 24 32 48 1000; do
  echo "== f64 low_lanes=$L max_cost=$C"; timeout 200 python bench.py $Q --option max_cost=$C --option low_lanes_fp64=$L 2>&1 | show
done; done
for L in 5 6 7; do for C in 32 48 1000; do
  echo "== f32 low_lanes=$L max_cost=$C"; timeout 200 python bench.py $Q --dtype f32 --option max_cost=$C --option low_lanes_fp32=$L 2>&1 | show
done; done
