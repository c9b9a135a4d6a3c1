#!/bin/bash
# round-1 GPU session G: re-measure the current tree (tests, bench lines, launch list, ncu full)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
( time timeout 900 python -m pytest tests -m gpu -q -x ) 2>&1 | tail -12 > gpurun_out/r1g_pytest_gpu.log
tail -6 gpurun_out/r1g_pytest_gpu.log
timeout 600 python bench.py > gpurun_out/r1g_bench_f64.json 2> gpurun_out/r1g_bench_f64.err; tail -c 1500 gpurun_out/r1g_bench_f64.json
timeout 600 python bench.py --dtype f32 > gpurun_out/r1g_bench_f32.json 2> gpurun_out/r1g_bench_f32.err; tail -c 1500 gpurun_out/r1g_bench_f32.json
timeout 300 python bench.py --impl reference > gpurun_out/r1g_bench_ref.json 2>&1; tail -c 600 gpurun_out/r1g_bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1g_launches.csv python bench.py --steps 2 --warmup 1 --depth 20 --no-e2e --no-cpu-baseline > gpurun_out/r1g_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tile_pass -s 6 -c 2 -o gpurun_out/r1g_tile_f64 -f python bench.py --qubits 28 --steps 1 --warmup 1 --depth 10 --no-e2e --no-cpu-baseline > gpurun_out/r1g_ncu_full.log 2>&1
tail -2 gpurun_out/r1g_ncu_full.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tile_pass -s 6 -c 2 -o gpurun_out/r1g_tile_f32 -f python bench.py --qubits 28 --dtype f32 --steps 1 --warmup 1 --depth 10 --no-e2e --no-cpu-baseline > gpurun_out/r1g_ncu_full32.log 2>&1
tail -2 gpurun_out/r1g_ncu_full32.log
