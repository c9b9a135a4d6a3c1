#!/bin/bash
# round-2 GPU session y: QFT-33 wall clock, lazy vs eager |0...0>, repeated
mkdir -p gpurun_out
timeout 600 python tools/qft_breakdown.py 33 4 2>&1 | tail -4
timeout 600 python tools/qft_breakdown.py 33 4 lazy_reset=0 2>&1 | tail -4
timeout 600 python run_configs.py qft --qubits 33 --repeat 3 2>&1 | tail -1 | cut -c1-400
