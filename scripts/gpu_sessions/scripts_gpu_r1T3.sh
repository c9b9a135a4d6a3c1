#!/bin/bash
# round-1 GPU session T3: complex128 artefacts with the final defaults (10-lane tiles, pass cost 24)
mkdir -p gpurun_out
( timeout 300 python -m pytest tests -m gpu -q ) 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/r1T3_bench_f64.json 2> gpurun_out/r1T3_bench_f64.err; tail -c 2700 gpurun_out/r1T3_bench_f64.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1T3_launches.csv python bench.py --steps 2 --warmup 1 --depth 20 --no-e2e --no-cpu-baseline > gpurun_out/r1T3_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tma_pass -s 6 -c 2 -o gpurun_out/r1T3_tma_f64_30q -f python bench.py --steps 1 --warmup 1 --depth 10 --no-e2e --no-cpu-baseline > gpurun_out/r1T3_ncu_f64.log 2>&1
tail -1 gpurun_out/r1T3_ncu_f64.log
