#!/bin/bash
# round-1 GPU session ZA: memory floor of a pass as a function of the contiguous run length (low lanes)
mkdir -p gpurun_out
Q="--steps 1 --warmup 1 --no-e2e --no-cpu-baseline --depth 6 --option tma_ws=0"
show() { python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()[:300]); continue
    r=d['roofline']; print('passes %d gates/pass %.1f avg_ms %.2f GB/s %.0f frac %.3f'%(r['launches_per_step'],r['gates_per_launch'],r['avg_launch_ms'],r['achieved'],r['frac']))
"; }
for L in 3 4 5 6 7 9; do
  echo "== f64 max_gates_per_pass=1 low_lanes=$L"; timeout 200 python bench.py $Q --option max_gates_per_pass=1 --option low_lanes_fp64=$L 2>&1 | show
done
for L in 4 5 6 8; do
  echo "== f32 max_gates_per_pass=1 low_lanes=$L"; timeout 200 python bench.py $Q --dtype f32 --option max_gates_per_pass=1 --option low_lanes_fp32=$L 2>&1 | show
done
