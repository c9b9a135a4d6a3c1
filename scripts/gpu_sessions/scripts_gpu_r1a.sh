#!/bin/bash
# round-1 GPU session A: parity suite, first bench lines, ncu launch list + full capture
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/r1a_pytest_gpu.log
cat gpurun_out/r1a_pytest_gpu.log
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/r1a_bench_f64.json 2> gpurun_out/r1a_bench_f64.err
tail -3 gpurun_out/r1a_bench_f64.err; cat gpurun_out/r1a_bench_f64.json
timeout 600 python bench.py --steps 3 --warmup 3 --dtype f32 --no-cpu-baseline > gpurun_out/r1a_bench_f32.json 2> gpurun_out/r1a_bench_f32.err
tail -3 gpurun_out/r1a_bench_f32.err; cat gpurun_out/r1a_bench_f32.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1a_launches.csv python bench.py --steps 2 --warmup 1 --depth 20 --no-e2e --no-cpu-baseline > gpurun_out/r1a_ncu_bench.log 2>&1
tail -2 gpurun_out/r1a_ncu_bench.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tile_pass -s 10 -c 2 -o gpurun_out/r1a_tile_f64 -f python bench.py --qubits 28 --steps 1 --warmup 1 --depth 10 --no-e2e --no-cpu-baseline > gpurun_out/r1a_ncu_full.log 2>&1
tail -2 gpurun_out/r1a_ncu_full.log
ls -la gpurun_out
