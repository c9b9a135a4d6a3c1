#!/bin/bash
# round-1 GPU session Q (4 GPUs): world-4 sharded parity (2-lane all-to-all exchange), config 4 on 4 GPUs, 4-GPU bench
mkdir -p gpurun_out
( time timeout 500 python -m pytest tests/test_dist_nccl.py -m gpu -q -x -k "random or grover" ) 2>&1 | tail -40 > gpurun_out/r1q_pytest_nccl4.log
tail -6 gpurun_out/r1q_pytest_nccl4.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512 run_configs.py pe --qubits 32 > gpurun_out/r1q_pe_4gpu.json 2> gpurun_out/r1q_pe_4gpu.err
tail -c 1300 gpurun_out/r1q_pe_4gpu.json; grep -i "error" gpurun_out/r1q_pe_4gpu.err | tail -3
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 4 --steps 2 --warmup 1 --depth 40 --no-cpu-baseline > gpurun_out/r1q_bench_4gpu.json 2> gpurun_out/r1q_bench_4gpu.err
tail -c 2500 gpurun_out/r1q_bench_4gpu.json; grep -i "error" gpurun_out/r1q_bench_4gpu.err | tail -3
