#!/bin/bash
# round-2 GPU session zc (2 GPUs): what sim.terminate() of a sharded simulator costs
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 tools/dist_terminate_breakdown.py 31 4 > gpurun_out/r2zc.log 2>&1
grep "^run\|Error\|error" gpurun_out/r2zc.log | head
