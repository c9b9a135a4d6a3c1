#!/bin/bash
# round-1 GPU session R (2 GPUs): commutation-aware exchange scheduling in dist.py
mkdir -p gpurun_out
( time timeout 400 python -m pytest tests/test_dist_nccl.py -m gpu -q -x -k "p2p" ) 2>&1 | tail -30 > gpurun_out/r1r_pytest_nccl.log
tail -5 gpurun_out/r1r_pytest_nccl.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 1 --depth 200 --no-cpu-baseline > gpurun_out/r1r_bench_2gpu.json 2> gpurun_out/r1r_bench_2gpu.err
tail -c 2600 gpurun_out/r1r_bench_2gpu.json; grep -i "error" gpurun_out/r1r_bench_2gpu.err | tail -3
