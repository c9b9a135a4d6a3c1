#!/bin/bash
# round-1 GPU session N (2 GPUs): sharded parity over NCCL / peer memory, 2-GPU bench, config 4
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
( time timeout 600 python -m pytest tests/test_dist_nccl.py -m gpu -q -x -k collective ) 2>&1 | tail -40 > gpurun_out/r1n_pytest_nccl_collective.log
tail -5 gpurun_out/r1n_pytest_nccl_collective.log
( time timeout 420 python -m pytest tests/test_dist_nccl.py -m gpu -q -x -k p2p ) 2>&1 | tail -60 > gpurun_out/r1n_pytest_nccl_p2p.log
tail -30 gpurun_out/r1n_pytest_nccl_p2p.log
for ex in collective p2p; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 1 --depth 40 --exchange $ex > gpurun_out/r1n_bench_2gpu_$ex.json 2> gpurun_out/r1n_bench_2gpu_$ex.err
tail -c 1500 gpurun_out/r1n_bench_2gpu_$ex.json; grep -i "error" gpurun_out/r1n_bench_2gpu_$ex.err | tail -3
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 run_configs.py pe --qubits 32 > gpurun_out/r1n_pe32_2gpu.json 2> gpurun_out/r1n_pe32_2gpu.err
tail -c 1200 gpurun_out/r1n_pe32_2gpu.json; grep -i "error" gpurun_out/r1n_pe32_2gpu.err | tail -3
