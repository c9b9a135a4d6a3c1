#!/bin/bash
# round-1 GPU session M: 16 amplitudes per thread for complex128 (reg_bits_fp64=4)
mkdir -p gpurun_out
Q="--steps 2 --warmup 1 --no-e2e --no-cpu-baseline --depth 60"
show() { python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()[:300]); continue
    r=d['roofline']; print('value %.3e ms/step %.0f passes %d gates/pass %.1f avg_ms %.2f GB/s %.0f frac %.3f p0 %.12f'%(d['value'],d['ms_per_step'],r['launches_per_step'],r['gates_per_launch'],r['avg_launch_ms'],r['achieved'],r['frac'],d['p0_check']))
"; }
for opt in "--option ctas_per_sm=2" "--option reg_bits_fp64=4" "--option reg_bits_fp64=4 --option tile_lanes_fp64=10" "--option reg_bits_fp64=4 --option tile_lanes_fp64=12" "--option reg_bits_fp64=4 --option tma_buffers=3" "--option reg_bits_fp64=4 --option ctas_per_sm=2" "--option reg_bits_fp64=4 --option max_gates_per_pass=12"; do
  echo "== f64 $opt"; timeout 300 python bench.py $Q $opt 2>&1 | show
done
