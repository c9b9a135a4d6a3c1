#!/bin/bash
# round-2 GPU session zb (2 GPUs, final code): the driver's bench command at N=2 with the final code (QFT-31 and QFT-34 in
# its qft sub-record).
mkdir -p gpurun_out
( time timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 5 --warmup 3 ) > gpurun_out/r2zb_bench_2gpu.json 2> gpurun_out/r2zb_bench_2gpu.err; tail -4 gpurun_out/r2zb_bench_2gpu.err
python - <<'PY'
import json
for tag in ('r2zb_bench_2gpu',):
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
