#!/bin/bash
# round-1 GPU session Z: warp-specialised pass kernel (producer warp owns the TMA traffic) — parity under a
# short timeout first (a deadlock must not hold the box), then A/B against the unspecialised kernel
mkdir -p gpurun_out
( time timeout 120 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "qft20 or large_state or fused_equals" ) 2>&1 | tail -6 > gpurun_out/r1z_pytest_first.log
tail -3 gpurun_out/r1z_pytest_first.log
if grep -q "passed" gpurun_out/r1z_pytest_first.log && ! grep -q "failed" gpurun_out/r1z_pytest_first.log; then
( time timeout 400 python -m pytest tests -m gpu -q -x ) 2>&1 | tail -6 > gpurun_out/r1z_pytest_gpu.log
tail -3 gpurun_out/r1z_pytest_gpu.log
Q="--steps 2 --warmup 1 --no-e2e --no-cpu-baseline --depth 60"
show() { python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()[:300]); continue
    r=d['roofline']; print('value %.3e ms/step %.0f passes %d gates/pass %.1f avg_ms %.2f GB/s %.0f frac %.3f p0 %.9f clk %s'%(d['value'],d['ms_per_step'],r['launches_per_step'],r['gates_per_launch'],r['avg_launch_ms'],r['achieved'],r['frac'],d['p0_check'],d['clocks']['sm_mhz']))
"; }
for opt in "--option tma_ws=0" "--option tma_ws=1" "--option tma_ws=1 --option max_cost=48" "--option tma_ws=1 --option tma_buffers=3"; do
  echo "== f64 $opt"; timeout 200 python bench.py $Q $opt 2>&1 | show
done
for opt in "--option tma_ws=0" "--option tma_ws=1" "--option tma_ws=1 --option tma_buffers=2"; do
  echo "== f32 $opt"; timeout 200 python bench.py $Q --dtype f32 $opt 2>&1 | show
done
fi
