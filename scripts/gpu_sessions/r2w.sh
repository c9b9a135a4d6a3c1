#!/bin/bash
# round-2 GPU session w: ncu evidence of the final code: launch list of the bench command (short), full
# captures of the fused pass in both precisions (kept small: one launch each, so both reports fit the
# 64 MiB return limit), tightened QFT tolerance tests.
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "large_state or fan or qft" ) > gpurun_out/r2w_pytest.log 2>&1; tail -3 gpurun_out/r2w_pytest.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2w_launches.csv python bench.py --steps 2 --warmup 1 --depth 20 --no-e2e --no-cpu-baseline --no-extras > gpurun_out/r2w_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tma_pass -s 6 -c 1 -o gpurun_out/r2w_tma_f64_30q -f python bench.py --steps 1 --warmup 1 --depth 10 --no-e2e --no-cpu-baseline --no-extras > gpurun_out/r2w_ncu_f64.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:tma_pass -s 6 -c 1 -o gpurun_out/r2w_tma_f32_30q -f python bench.py --dtype f32 --steps 1 --warmup 1 --depth 10 --no-e2e --no-cpu-baseline --no-extras > gpurun_out/r2w_ncu_f32.log 2>&1
tail -2 gpurun_out/r2w_ncu_f64.log; ls -la gpurun_out/
