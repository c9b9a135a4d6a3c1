#!/bin/bash
# round-2 GPU session k: wall-clock breakdown of a QFT run (front end / initialisation / passes) at 30 and
# 33 qubits; one full ncu capture of a QFT-30 fan pass.
mkdir -p gpurun_out
timeout 300 python tools/qft_breakdown.py 30 3 2>&1 | tail -4
timeout 300 python tools/qft_breakdown.py 30 2 fan_cost=1 2>&1 | tail -2
timeout 600 python tools/qft_breakdown.py 33 3 2>&1 | tail -4
timeout 600 python tools/qft_breakdown.py 33 2 fan_cost=1 2>&1 | tail -2
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tma_pass -s 1 -c 1 -o gpurun_out/r2k_qft30_pass -f python run_configs.py qft --qubits 30 > gpurun_out/r2k_ncu_qft30.log 2>&1
tail -2 gpurun_out/r2k_ncu_qft30.log
