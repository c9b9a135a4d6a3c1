#!/bin/bash
# round-1 GPU session T2 (final code): the judged artefacts — full GPU suite, smoke, bench lines (both dtypes + reference arm),
# ncu launch list and full captures at the bench size
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q ) 2>&1 | tail -8 > gpurun_out/r1T2_pytest_gpu.log
tail -4 gpurun_out/r1T2_pytest_gpu.log
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/r1T2_bench_f64.json 2> gpurun_out/r1T2_bench_f64.err; tail -c 2600 gpurun_out/r1T2_bench_f64.json
timeout 600 python bench.py --dtype f32 > gpurun_out/r1T2_bench_f32.json 2> gpurun_out/r1T2_bench_f32.err; tail -c 2600 gpurun_out/r1T2_bench_f32.json
timeout 300 python bench.py --impl reference > gpurun_out/r1T2_bench_ref.json 2>&1; tail -c 500 gpurun_out/r1T2_bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1T2_launches.csv python bench.py --steps 2 --warmup 1 --depth 20 --no-e2e --no-cpu-baseline > gpurun_out/r1T2_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tma_pass -s 6 -c 2 -o gpurun_out/r1T2_tma_f64_30q -f python bench.py --steps 1 --warmup 1 --depth 10 --no-e2e --no-cpu-baseline > gpurun_out/r1T2_ncu_f64.log 2>&1
tail -1 gpurun_out/r1T2_ncu_f64.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tma_pass -s 6 -c 2 -o gpurun_out/r1T2_tma_f32_30q -f python bench.py --dtype f32 --steps 1 --warmup 1 --depth 10 --no-e2e --no-cpu-baseline > gpurun_out/r1T2_ncu_f32.log 2>&1
tail -1 gpurun_out/r1T2_ncu_f32.log
