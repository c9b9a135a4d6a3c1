#!/bin/bash
# round-1 GPU session P (2 GPUs): p2p parity after the handle-exchange race fix, configs 4 / 5 at 2 GPUs
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_dist_nccl.py -m gpu -q -x ) 2>&1 | tail -40 > gpurun_out/r1p_pytest_nccl.log
tail -6 gpurun_out/r1p_pytest_nccl.log
for cfg in "pe --qubits 32" "qft --qubits 33" "grover --qubits 31"; do
name=$(echo $cfg | cut -d' ' -f1)
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 run_configs.py $cfg > gpurun_out/r1p_${name}_2gpu.json 2> gpurun_out/r1p_${name}_2gpu.err
tail -c 1300 gpurun_out/r1p_${name}_2gpu.json; grep -i "error" gpurun_out/r1p_${name}_2gpu.err | tail -3
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 1 --depth 40 > gpurun_out/r1p_bench_2gpu.json 2> gpurun_out/r1p_bench_2gpu.err
tail -c 1500 gpurun_out/r1p_bench_2gpu.json; grep -i "error" gpurun_out/r1p_bench_2gpu.err | tail -3
