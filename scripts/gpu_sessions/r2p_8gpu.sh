#!/bin/bash
# round-2 GPU session p (8 GPUs): the driver's bench command at N=8 (QFT-33 and QFT-35 in its qft sub-record),
# then the in-place exchange for comparison.
mkdir -p gpurun_out
( time timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 5 --warmup 3 ) > gpurun_out/r2p_bench_8gpu.json 2> gpurun_out/r2p_bench_8gpu.err; tail -4 gpurun_out/r2p_bench_8gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 8 --steps 3 --warmup 3 --exchange p2p_inplace --no-extras --no-e2e > gpurun_out/r2p_bench_8gpu_inplace.json 2> gpurun_out/r2p_bench_8gpu_inplace.err
python - <<'PY'
import json
for tag in ('r2p_bench_8gpu', 'r2p_bench_8gpu_inplace'):
    try:
        d = json.loads(open('gpurun_out/%s.json' % tag).read().strip().splitlines()[-1])
        r, nv = d['roofline'], d['nvlink']
        print(tag, 'upd/s %.3e ms/step %.1f frac %.3f passes %.0f | e2e %s | nvlink %s | f32 %s | qft %s' % (
            d['value'], d['ms_per_step'], r['frac'], r['launches_per_step'], d['e2e'] and '%.3e' % d['e2e']['value'],
            {k: nv[k] for k in ('exchange', 'exchanges_per_step', 'lanes_per_exchange', 'ms_per_step', 'achieved', 'frac')},
            d.get('f32') and '%.3e' % d['f32']['value'], d.get('qft')))
    except Exception as e:
        print(tag, 'failed', e, open('gpurun_out/%s.err' % tag).read()[-1200:])
PY
