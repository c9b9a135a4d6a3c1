#!/bin/bash
# round-2 GPU session ze (2 GPUs): kept IPC mappings + exported blocks pinned in the pool: the size
# sequence that ran out of memory before, the NCCL parity suite, the bench command with sub-records
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 tools/dist_mem_sequence.py 31 31 34 31 > gpurun_out/r2ze_seq.log 2>&1
grep "QFT-\|FAILED\|Out of" gpurun_out/r2ze_seq.log | head -12
( time timeout 900 python -m pytest tests/test_dist_nccl.py -m gpu -q -x ) > gpurun_out/r2ze_pytest_nccl.log 2>&1; tail -3 gpurun_out/r2ze_pytest_nccl.log
( time timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 5 --warmup 3 ) > gpurun_out/r2ze_bench_2gpu.json 2> gpurun_out/r2ze_bench_2gpu.err; tail -3 gpurun_out/r2ze_bench_2gpu.err
python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/r2ze_bench_2gpu.json').read().strip().splitlines()[-1])
    r, nv = d['roofline'], d['nvlink']
    print('upd/s %.3e ms/step %.1f frac %.3f | e2e %.3e split %s | nvlink %.0f GB/s | f32 %.3e | qft %s' % (
        d['value'], d['ms_per_step'], r['frac'], d['e2e']['value'], d['e2e']['split_ms'], nv['achieved'], d['f32']['value'],
        [(q['qubits'], round(q.get('ms', -1), 1), round(q.get('first_run_ms', -1), 1), q.get('ok'), q.get('error')) for q in d['qft']]))
except Exception as e:
    print('failed', e, open('gpurun_out/r2ze_bench_2gpu.err').read()[-1200:])
PY
