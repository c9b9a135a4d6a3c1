#!/bin/bash
# round-1 GPU session U (8 GPUs): config 5 (QFT 35 qubits, 512 GiB) and the 8-GPU bench line
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,memory.total --format=csv,noheader | head -8
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 run_configs.py qft --qubits 35 > gpurun_out/r1u_qft35_8gpu.json 2> gpurun_out/r1u_qft35_8gpu.err
tail -c 1300 gpurun_out/r1u_qft35_8gpu.json; grep -i "error" gpurun_out/r1u_qft35_8gpu.err | tail -3
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r1u_bench_8gpu.json 2> gpurun_out/r1u_bench_8gpu.err
rc=$?
tail -c 2800 gpurun_out/r1u_bench_8gpu.json; grep -i "error" gpurun_out/r1u_bench_8gpu.err | tail -3
if [ $rc -ne 0 ]; then
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --steps 2 --warmup 1 --no-cpu-baseline --exchange collective > gpurun_out/r1u_bench_8gpu_collective.json 2> gpurun_out/r1u_bench_8gpu_collective.err
tail -c 2800 gpurun_out/r1u_bench_8gpu_collective.json; grep -i "error" gpurun_out/r1u_bench_8gpu_collective.err | tail -3
fi
