#!/bin/bash
# round-2 GPU session zd (2 GPUs): C-ABI call profile of the bench's e2e leg (what makes terminate slow)
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29581 tools/capi_profile.py bench.py --gpus 2 --steps 3 --warmup 1 --no-extras > gpurun_out/r2zd.json 2> gpurun_out/r2zd.err
grep "capi\]" gpurun_out/r2zd.err | head -14
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2zd.json').read().strip().splitlines()[-1])
print(d['e2e']['split_ms'], d['e2e']['ms_per_step'], d['ms_per_step'])
PY
