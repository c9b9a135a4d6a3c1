#!/bin/bash
# round-1 GPU session O: packed FFMA2 complex64 path + K=4 complex128 default — parity, bench, ncu
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q -x ) 2>&1 | tail -25 > gpurun_out/r1o_pytest_gpu.log
tail -8 gpurun_out/r1o_pytest_gpu.log
Q="--steps 2 --warmup 1 --no-e2e --no-cpu-baseline --depth 60"
show() { python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()[:300]); continue
    r=d['roofline']; print('value %.3e ms/step %.0f passes %d gates/pass %.1f avg_ms %.2f GB/s %.0f frac %.3f p0 %.9f'%(d['value'],d['ms_per_step'],r['launches_per_step'],r['gates_per_launch'],r['avg_launch_ms'],r['achieved'],r['frac'],d['p0_check']))
"; }
for opt in "" ; do
  echo "== f64 $opt"; timeout 300 python bench.py $Q $opt 2>&1 | show
done
for opt in "" "--option ctas_per_sm=2" "--option tile_lanes_fp32=11" "--option tile_lanes_fp32=13" "--option tma_buffers=3"; do
  echo "== f32 $opt"; timeout 300 python bench.py $Q --dtype f32 $opt 2>&1 | show
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tma_pass -s 6 -c 2 -o gpurun_out/r1o_tma_f32 -f python bench.py --qubits 28 --dtype f32 --steps 1 --warmup 1 --depth 10 --no-e2e --no-cpu-baseline > gpurun_out/r1o_ncu_full32.log 2>&1
tail -2 gpurun_out/r1o_ncu_full32.log
timeout 300 python run_configs.py grover --qubits 30 > gpurun_out/r1o_grover30.json 2> gpurun_out/r1o_grover30.err; tail -c 1500 gpurun_out/r1o_grover30.json; tail -3 gpurun_out/r1o_grover30.err
