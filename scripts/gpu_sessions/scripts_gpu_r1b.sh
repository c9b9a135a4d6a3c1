#!/bin/bash
# round-1 GPU session B: new op bodies — parity, bench, option sweep, ncu
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/r1b_pytest_gpu.log
cat gpurun_out/r1b_pytest_gpu.log
Q="--steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
for opt in "" "--option tile_lanes_fp64=12" "--option tile_lanes_fp64=13" "--option low_lanes_fp64=3" "--option low_lanes_fp64=4 --option tile_lanes_fp64=12" "--option max_gates_per_pass=12" "--option max_gates_per_pass=16 --option tile_lanes_fp64=12"; do
  echo "== f64 $opt"
  timeout 300 python bench.py $Q $opt 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()[:300]); continue
    r=d['roofline']; print('value %.3e ms/step %.0f passes %d gates/pass %.1f avg_ms %.2f GB/s %.0f frac %.3f'%(d['value'],d['ms_per_step'],r['launches_per_step'],r['gates_per_launch'],r['avg_launch_ms'],r['achieved'],r['frac']))
"
done
for opt in "" "--option tile_lanes_fp32=13" "--option tile_lanes_fp32=14" "--option low_lanes_fp32=4"; do
  echo "== f32 $opt"
  timeout 300 python bench.py $Q --dtype f32 $opt 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()[:300]); continue
    r=d['roofline']; print('value %.3e ms/step %.0f passes %d gates/pass %.1f avg_ms %.2f GB/s %.0f frac %.3f'%(d['value'],d['ms_per_step'],r['launches_per_step'],r['gates_per_launch'],r['avg_launch_ms'],r['achieved'],r['frac']))
"
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tile_pass -s 10 -c 2 -o gpurun_out/r1b_tile_f64 -f python bench.py --qubits 28 --steps 1 --warmup 1 --depth 10 --no-e2e --no-cpu-baseline > gpurun_out/r1b_ncu_full.log 2>&1
tail -2 gpurun_out/r1b_ncu_full.log
