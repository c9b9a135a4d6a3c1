#!/bin/bash
# round-2 GPU session a: power / clock model of the fused pass (what bounds a SUSTAINED pass) with the
# round-1 kernel: sustained copy, memory-floor pass, 6 / 8 / 12 dense ops per pass.
mkdir -p gpurun_out
python - <<'PY' > gpurun_out/r2a_copy_power.json 2> gpurun_out/r2a_copy_power.err
import json, subprocess, threading, time, torch
samples = []
p = subprocess.Popen(['nvidia-smi', '--query-gpu=clocks.sm,power.draw,clocks.mem', '--format=csv,noheader,nounits', '-lms', '100'],
                     stdout=subprocess.PIPE, text=True)
def rd():
    for line in p.stdout:
        samples.append((time.time(), line.strip()))
threading.Thread(target=rd, daemon=True).start()
a = torch.empty(1 << 30, dtype=torch.float64, device='cuda')   # 8 GiB
b = torch.empty_like(a)
torch.cuda.synchronize()
out = {}
for label, secs in (('burst', 0.3), ('sustained', 6.0)):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 0
    t0 = time.time()
    e0.record()
    while time.time() - t0 < secs:
        for _ in range(10):
            b.copy_(a)
        n += 10
        torch.cuda.synchronize()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    tail = [s for t, s in samples if t >= t0 + secs * 0.5]
    out[label] = {'gbs': 2 * a.numel() * 8 * n / (ms * 1e-3) / 1e9, 'copies': n, 'smi_tail': tail[-5:]}
p.terminate()
print(json.dumps(out))
PY
cat gpurun_out/r2a_copy_power.json
for cfg in "max_gates_per_pass=1" "max_cost=16" "max_cost=24" "max_cost=32" "max_cost=48" "max_cost=48 --option tile_lanes_fp64=11"; do
  tag=$(echo "$cfg" | tr ' =-' '___')
  timeout 400 python bench.py --steps 3 --warmup 3 --depth 60 --no-e2e --no-cpu-baseline --option $cfg > gpurun_out/r2a_$tag.json 2> gpurun_out/r2a_$tag.err
  python - "$cfg" gpurun_out/r2a_$tag.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    r = d['roofline']
    print(sys.argv[1], 'ms/pass %.3f frac %.3f upd/s %.3e passes %.0f sm %s W %s %s' % (r['avg_launch_ms'], r['frac'], d['value'], r['launches_per_step'], d['clocks']['sm_mhz'], d['clocks'].get('power_w_max'), d['clocks']['reasons']))
except Exception as e:
    print(sys.argv[1], 'failed', e)
PY
done
