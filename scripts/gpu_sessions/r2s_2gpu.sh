#!/bin/bash
# round-2 GPU session s (2 GPUs): sharded QFT breakdown with retained IPC mappings and the relaxed pool cache; NCCL suite
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 tools/dist_qft_breakdown.py 31 3 2>&1 | grep -v "^\*\|OMP_NUM" | tail -5
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29552 tools/dist_qft_breakdown.py 34 3 2>&1 | grep -v "^\*\|OMP_NUM" | tail -5
( time timeout 900 python -m pytest tests/test_dist_nccl.py -m gpu -q -x ) > gpurun_out/r2s_pytest_nccl.log 2>&1; tail -4 gpurun_out/r2s_pytest_nccl.log
