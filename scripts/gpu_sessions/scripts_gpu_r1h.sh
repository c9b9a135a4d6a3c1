#!/bin/bash
# round-1 GPU session H: full GPU parity suite after the simple-path multiplexer fix
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -q ) 2>&1 | tail -25 > gpurun_out/r1h_pytest_gpu.log
tail -8 gpurun_out/r1h_pytest_gpu.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3
