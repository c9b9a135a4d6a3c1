#!/bin/bash
# round-1 GPU session I (2 GPUs): sharded parity over NCCL / peer memory + a 2-GPU bench line
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
( time timeout 900 python -m pytest tests/test_dist_nccl.py -m gpu -q -x ) 2>&1 | tail -25 > gpurun_out/r1i_pytest_nccl.log
tail -8 gpurun_out/r1i_pytest_nccl.log
for ex in p2p collective; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 1 --depth 40 --exchange $ex > gpurun_out/r1i_bench_2gpu_$ex.json 2> gpurun_out/r1i_bench_2gpu_$ex.err
tail -c 1800 gpurun_out/r1i_bench_2gpu_$ex.json; tail -3 gpurun_out/r1i_bench_2gpu_$ex.err
done
