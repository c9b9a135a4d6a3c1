#!/bin/bash
# round-1 GPU session ZC (2 GPUs): the driver's scaling command with the final code
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r1zc_bench_2gpu.json 2> gpurun_out/r1zc_bench_2gpu.err
tail -c 2700 gpurun_out/r1zc_bench_2gpu.json; grep -i "error" gpurun_out/r1zc_bench_2gpu.err | tail -3
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 2>&1 | tail -c 700
