#!/bin/bash
# round-2 GPU session za (2 GPUs): sharded states of different sizes in one process (memory hand-over)
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 tools/dist_mem_sequence.py 31 31 34 31 > gpurun_out/r2za.log 2>&1
grep -v "^\*\|OMP_NUM" gpurun_out/r2za.log | grep -B2 -A25 "QFT-\|FAILED" | head -70
