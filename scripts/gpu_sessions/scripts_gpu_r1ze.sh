#!/bin/bash
# round-1 GPU session ZE: complex64 pass cost caps
mkdir -p gpurun_out
Q="--steps 1 --warmup 1 --no-e2e --no-cpu-baseline --depth 60 --dtype f32"
show() { python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()[:300]); continue
    r=d['roofline']; print('value %.3e passes %d gates/pass %.1f avg_ms %.2f GB/s %.0f frac %.3f clk %s'%(d['value'],r['launches_per_step'],r['gates_per_launch'],r['avg_launch_ms'],r['achieved'],r['frac'],d['clocks']['sm_mhz']))
"; }
for opt in "--option max_cost=16" "--option max_cost=20" "--option max_cost=24" "--option max_cost=24 --option tile_lanes_fp32=11" "--option max_cost=20 --option tile_lanes_fp32=11"; do
  echo "== f32 $opt"; timeout 100 python bench.py $Q $opt 2>&1 | show
done
