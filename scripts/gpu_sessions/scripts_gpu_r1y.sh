#!/bin/bash
# round-1 GPU session Y: shared-memory slots recomputed at the stage store (register pressure)
mkdir -p gpurun_out
( time timeout 300 python -m pytest tests -m gpu -q -x ) 2>&1 | tail -5 > gpurun_out/r1y_pytest_gpu.log
head -1 gpurun_out/r1y_pytest_gpu.log
Q="--steps 2 --warmup 1 --no-e2e --no-cpu-baseline --depth 60"
show() { python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()[:300]); continue
    r=d['roofline']; print('value %.3e ms/step %.0f passes %d gates/pass %.1f avg_ms %.2f GB/s %.0f frac %.3f p0 %.9f clk %s'%(d['value'],d['ms_per_step'],r['launches_per_step'],r['gates_per_launch'],r['avg_launch_ms'],r['achieved'],r['frac'],d['p0_check'],d['clocks']['sm_mhz']))
"; }
for opt in "" ; do
  echo "== f64 $opt"; timeout 300 python bench.py $Q $opt 2>&1 | show
done
for opt in "" "--option tma_buffers=2" ; do
  echo "== f32 $opt"; timeout 300 python bench.py $Q --dtype f32 $opt 2>&1 | show
done
