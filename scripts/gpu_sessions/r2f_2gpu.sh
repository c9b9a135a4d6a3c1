#!/bin/bash
# round-2 GPU session f (2 GPUs): sharded parity over NCCL / NVLink (push, in-place, collective),
# bench at N=2 with the push and the in-place exchange.
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_dist_nccl.py -m gpu -q -x ) > gpurun_out/r2f_pytest_nccl.log 2>&1; tail -8 gpurun_out/r2f_pytest_nccl.log
for ex in p2p p2p_inplace; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --exchange $ex --no-extras > gpurun_out/r2f_bench_2gpu_$ex.json 2> gpurun_out/r2f_bench_2gpu_$ex.err
  python - $ex <<'PY'
import json, sys
ex = sys.argv[1]
try:
    d = json.loads(open('gpurun_out/r2f_bench_2gpu_%s.json' % ex).read().strip().splitlines()[-1])
    r, nv = d['roofline'], d['nvlink']
    print(ex, 'upd/s %.3e ms/step %.1f frac %.3f passes %.0f | e2e %.3e | nvlink %s' % (d['value'], d['ms_per_step'], r['frac'], r['launches_per_step'], d['e2e']['value'], {k: nv[k] for k in ('exchange', 'exchanges_per_step', 'lanes_per_exchange', 'ms_per_step', 'achieved', 'frac')}))
except Exception as e:
    print(ex, 'failed', e, open('gpurun_out/r2f_bench_2gpu_%s.err' % ex).read()[-800:])
PY
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 2 --warmup 1 --depth 20 --no-e2e > gpurun_out/r2f_bench_2gpu_extras.json 2> gpurun_out/r2f_bench_2gpu_extras.err; python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/r2f_bench_2gpu_extras.json').read().strip().splitlines()[-1])
    print('extras at N=2: f32 %.3e | qft %s' % (d['f32']['value'], d['qft']))
except Exception as e:
    print('extras failed', e, open('gpurun_out/r2f_bench_2gpu_extras.err').read()[-800:])
PY
