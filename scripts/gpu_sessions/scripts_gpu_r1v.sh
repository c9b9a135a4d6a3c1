#!/bin/bash
# round-1 GPU session V (4 GPUs): full sharded parity suite (joins, unrolled exchange kernel) + 4-GPU bench with the exchange scheduler
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_dist_nccl.py -m gpu -q ) 2>&1 | tail -30 > gpurun_out/r1v_pytest_nccl.log
tail -6 gpurun_out/r1v_pytest_nccl.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 4 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r1v_bench_4gpu.json 2> gpurun_out/r1v_bench_4gpu.err
tail -c 2800 gpurun_out/r1v_bench_4gpu.json; grep -i "error" gpurun_out/r1v_bench_4gpu.err | tail -3
