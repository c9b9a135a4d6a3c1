#!/usr/bin/env python
"""bench.py — the hot path measured on B200 (contract: see DESIGN.md §Measurement).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--qubits 30] [--depth 200] [--dtype f64|f32]

Workload (BASELINE.json configs[1] generator at the qubit count the metric is quoted on):
the seeded random circuit of qgate_b200/circuits.py:random_u3_cx — per layer a U3 with
uniform random angles on every qubit, then a CX ladder — `depth` layers on `qubits`
qubits, complex128 unless --dtype f32.  One STEP = the whole circuit applied once to a
freshly reset state vector.  At N GPUs the state has qubits + log2(N) qubits, sharded on the
high-order (global) qubits, one shard per rank (weak scaling: the shard size is fixed).

  value  gate-amplitude updates / s = sum over gates of 2^(n - n_controls) / device time,
         gates already queued in the engine and the state resident in HBM; device time from
         CUDA events on the engine's stream, max over ranks.
  e2e    the same metric through the public API (qgate_b200.simulator.cuda().run(circuit) on
         host objects, then calc_probability + a states[] slice read back to NumPy), wall clock
         around the call with the device drained on both sides.
  roofline  the fused tile pass: algorithmic bytes per launch = 2 * 2^n_local * sizeof(complex)
         divided by the average pass duration inside the timed region, against the measured
         copy bandwidth of MEASURED_PEAKS.json.
  cpu_baseline  the reference's own CPU runtime (oracle/_ref, compiled unmodified from
         /root/reference) on the host cores, on a bounded sample of the same workload.

--impl reference times that CPU runtime alone and prints the same line with "impl": "reference".
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

import numpy as np  # noqa: E402

METRIC = 'gate-amp updates/s'
UNIT = 'updates/s'


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=3)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=('ours', 'reference'))
    ap.add_argument('--qubits', type=int, default=30, help='qubits per GPU shard x 1 GPU')
    ap.add_argument('--depth', type=int, default=200)
    ap.add_argument('--dtype', default='f64', choices=('f64', 'f32'))
    ap.add_argument('--seed', type=int, default=1234)
    ap.add_argument('--cpu-sample-gates', type=int, default=45,
                    help='gates of layer 0 the CPU reference runs per step (45 = the whole layer at 30 qubits, ~12 s)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-extras', action='store_true',
                    help='skip the sub-records (the other precision, the QFT runs)')
    ap.add_argument('--own-frontend', action='store_true',
                    help="e2e through this package's front end even when the reference's is installed")
    ap.add_argument('--option', action='append', default=[], help='engine option name=value')
    ap.add_argument('--exchange', default='auto', choices=('auto', 'p2p', 'p2p_inplace', 'collective'),
                    help='lane exchange engine of the sharded runs (p2p pushes into spare buffers when '
                         'memory allows; p2p_inplace forces the in-place swap kernel)')
    return ap.parse_args()


# ---- the workload as plain tuples (so both arms and both layers see identical gates) ----------

def random_circuit_gates(n, depth, seed):
    """[(u3 angles or None, control lane or -1, target lane)] in the order of
    circuits.random_u3_cx (same RandomState stream)."""
    rng = np.random.RandomState(seed)
    gates = []
    for d in range(depth):
        for i in range(n):
            theta, phi, lam = rng.uniform(0., 2. * math.pi, 3)
            gates.append(((float(theta), float(phi), float(lam)), -1, i))
        for i in range(d % 2, n - 1, 2):
            gates.append((None, i, i + 1))
    return gates


def updates_of(gates, n):
    return sum((1 << n) if g[1] < 0 else (1 << (n - 1)) for g in gates)


# ---- clocks -----------------------------------------------------------------------------------

class ClockSampler:
    FIELDS = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
              'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
              'clocks_event_reasons.sw_power_cap')

    def __init__(self, index=0):
        self.samples = []
        self.proc = None
        self.index = index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.FIELDS,
                 '--format=csv,noheader,nounits', '-lms', '200'],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            parts = [p.strip() for p in line.split(',')]
            if len(parts) >= 7:
                self.samples.append(parts)

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, power, reasons = [], [], [], set()
        names = ('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap')
        for parts in self.samples:
            try:
                sm.append(float(parts[0]))
                smax.append(float(parts[1]))
                power.append(float(parts[2]))
            except ValueError:
                continue
            for name, flag in zip(names, parts[3:7]):
                if flag.lower().startswith('active'):
                    reasons.add(name)
        # "under load": samples in the upper half of the power range
        load = [s for s, p in zip(sm, power) if power and p >= 0.5 * max(power)] or sm
        return {'sm_mhz': float(np.median(load)) if load else None,
                'sm_max_mhz': max(smax) if smax else None,
                'power_w_max': max(power) if power else None,
                'samples': len(sm), 'reasons': sorted(reasons)}


def measured_peak():
    path = os.path.join(REPO, 'MEASURED_PEAKS.json')
    try:
        with open(path) as f:
            return float(json.load(f)['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    except Exception:
        return 6650., 'fallback (B200_PROFILING.md)'


def measured_traffic(dtype, qubits, kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the
    committed `ncu --set full` capture of this workload (profiles/traffic.json); None when no
    capture matches the dtype / shard size / kernel."""
    try:
        with open(os.path.join(REPO, 'profiles', 'traffic.json')) as f:
            table = json.load(f)
        entry = table['{}_{}q'.format(dtype, qubits)]
        return entry['dram_bytes_per_launch'] if entry['kernel'] == kernel else None
    except Exception:
        return None


# ---- reference CPU runtime (oracle/_ref): cpu_baseline leg and --impl reference ---------------

class ReferenceCpu:
    """The unmodified reference CPU runtime behind the C ABI, driven gate by gate."""

    def __init__(self, n, dtype):
        from oracle import ref_runtime
        if not ref_runtime.available():
            ref_runtime.build()
        if not ref_runtime.available():
            raise RuntimeError('oracle/_ref/libqgate_ref_cpu.so is missing')
        self.module = ref_runtime.module
        self.n = n
        self.qstates = self.module.create_qubit_states(dtype)
        self.proc = self.qstates.processor
        self.proc.initialize_qubit_states(self.qstates, n)
        self.proc.reset_qubit_states(self.qstates)
        self.api = self.module.api

    def apply(self, gates):
        import ctypes as C
        from qgate_b200 import _capi
        lib = self.api.lib
        for angles, ctrl, target in gates:
            if ctrl < 0:
                args = (C.c_double * 3)(*angles)
                rc = lib.qgb_qproc_apply_gate_typed(self.proc.ptr, _capi.GATE_IDS['U'], args, 3, 0,
                                                    self.qstates.ptr, None, 0, target)
            else:
                args = (C.c_double * 1)()
                c = (C.c_int * 1)(ctrl)
                rc = lib.qgb_qproc_apply_gate_typed(self.proc.ptr, _capi.GATE_IDS['X'], args, 0, 0,
                                                    self.qstates.ptr, c, 1, target)
            self.api.check(rc)
        self.proc.synchronize()

    def close(self):
        self.qstates.delete()


def cpu_sample(gates, n, count):
    """A bounded sample of the workload: `count` gates spread evenly over layer 0 (U3 gates on
    low, middle and high lanes plus CX gates), at the full qubit count."""
    layer0 = gates[: n + len(range(0, n - 1, 2))]
    idx = np.unique(np.linspace(0, len(layer0) - 1, count).round().astype(int))
    return [layer0[i] for i in idx]


def cores_used():
    env = os.environ.get('QGATE_NUM_WORKERS')
    if env:
        return int(env)
    try:
        return len(os.sched_getaffinity(0))   # Parallel.cpp:40-45
    except AttributeError:
        return os.cpu_count()


def run_reference_arm(args, dtype, n):
    """--impl reference: the reference CPU runtime, all host threads, bounded sample per step."""
    gates = random_circuit_gates(n, 1, args.seed)
    sample = cpu_sample(gates, n, args.cpu_sample_gates)
    upd = updates_of(sample, n)
    ref = ReferenceCpu(n, dtype)
    for _ in range(args.warmup):
        ref.apply(sample)
    times = []
    for _ in range(args.steps):
        t0 = time.perf_counter()
        ref.apply(sample)
        times.append(time.perf_counter() - t0)
    ref.close()
    total = sum(times)
    value = upd * args.steps / total
    desc = '{} of the {} gates of layer 0 (U3 on low/mid/high lanes + CX) at {} qubits'.format(
        len(sample), len(gates), n)
    return value, 1e3 * total / args.steps, desc


# ---- our arm ------------------------------------------------------------------------------------

def submit_native(proc, qstates, gates):
    """Queue gates in the engine (deferred; nothing runs until flush)."""
    import ctypes as C
    from qgate_b200 import _capi
    lib, check = proc.api.lib, proc.api.check
    uid, xid = _capi.GATE_IDS['U'], _capi.GATE_IDS['X']
    noargs = (C.c_double * 1)()
    for angles, ctrl, target in gates:
        if ctrl < 0:
            rc = lib.qgb_qproc_apply_gate_typed(proc.ptr, uid, (C.c_double * 3)(*angles), 3, 0,
                                                qstates.ptr, None, 0, target)
        else:
            rc = lib.qgb_qproc_apply_gate_typed(proc.ptr, xid, noargs, 0, 0, qstates.ptr,
                                                (C.c_int * 1)(ctrl), 1, target)
        if rc:
            check(rc)


# FP64 / FP32 issue ceilings of the dense-gate arithmetic (tools/dfma_probe.cu on B200: one DFMA warp
# instruction per 2.1 cycles per SM partition; packed FFMA2: one per cycle), at the maximum SM clock
def pipe_roofline(dtype_name, stats, n_local, seconds, sm_max_mhz):
    sm_count, partitions = 148, 4
    clock = (sm_max_mhz or 1965.) * 1e6
    if dtype_name == 'f64':
        per_amp_shear, per_amp_direct, rate, unit = 6., 8., 32. / 2.1, 'DFMA/s'
    else:
        per_amp_shear, per_amp_direct, rate, unit = 3., 4., 32., 'FFMA2/s'
    instr = (per_amp_shear * stats.get('shear_ops', 0) + per_amp_direct * stats.get('direct_ops', 0)) \
        * float(1 << n_local)
    peak = sm_count * partitions * rate * clock
    achieved = instr / seconds if seconds > 0 else 0.
    return {'achieved': achieved, 'peak': peak, 'unit': unit, 'frac': achieved / peak,
            'per_amplitude': {'sheared 2x2': per_amp_shear, 'direct 2x2': per_amp_direct},
            'sheared_ops': stats.get('shear_ops', 0), 'direct_ops': stats.get('direct_ops', 0),
            'peak_source': 'tools/dfma_probe.cu issue rate x {:.0f} MHz'.format(clock / 1e6)}


def device_leg(args, dtype_name, n, gates, runtime, api, torch, dist, distributed, stream, steps, warmup,
               sample_clocks, local_rank):
    """gates queued in the engine, CUDA events around the flush; returns the numbers of one dtype"""
    dtype = np.float64 if dtype_name == 'f64' else np.float32
    elem = 16 if dtype_name == 'f64' else 8
    upd = updates_of(gates, n)
    qstates = runtime.create_qubit_states(dtype)
    proc = qstates.processor
    proc.initialize_qubit_states(qstates, n)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # this leg times the device alone: every gate is queued BEFORE the first event, so the engine must not
    # start launching passes while the queue fills (its default; the e2e leg below runs with the default)
    api.set_option('queue_stream', 0)

    def barrier():
        torch.cuda.synchronize()
        if distributed:
            dist.barrier()
            torch.cuda.synchronize()

    def device_step():
        proc.reset_qubit_states(qstates)
        if distributed:
            proc.submit_tuples(qstates, gates)
        else:
            submit_native(proc, qstates, gates)

    for _ in range(warmup):
        device_step()
        proc.flush(qstates)
    barrier()
    api.stats_reset()
    if distributed:
        runtime.ctx.exchange_ms()
        for key in runtime.ctx.stats:
            runtime.ctx.stats[key] = 0
    clocks = ClockSampler(local_rank) if sample_clocks else None
    if clocks:
        clocks.start()
    device_ms = 0.
    for _ in range(steps):
        device_step()
        barrier()
        with torch.cuda.stream(stream):
            ev0.record(stream)
            proc.flush(qstates)
            ev1.record(stream)
        barrier()
        device_ms += ev0.elapsed_time(ev1)
    clock_info = clocks.stop() if clocks else None
    stats = api.stats()
    p0 = proc.calc_probability(qstates, 0)     # keeps the result observable
    if distributed:
        t = torch.tensor([device_ms], dtype=torch.float64, device='cuda')
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        device_ms = float(t.item())
    exch_ms = runtime.ctx.exchange_ms() if distributed else 0.
    dstats = dict(runtime.ctx.stats) if distributed else None
    qstates.delete()
    api.set_option('queue_stream', 1)
    del qstates, proc
    n_local = n - int(round(math.log2(max(1, int(os.environ.get('WORLD_SIZE', '1'))))))
    tile_passes = stats['tile_passes']
    pass_bytes = 2 * (elem << n_local)
    avg_pass_ms = (device_ms - exch_ms) / max(1, tile_passes)
    peak, peak_src = measured_peak()
    achieved = pass_bytes / (avg_pass_ms * 1e-3) / 1e9
    kernel_name = 'tma_pass_kernel'
    pipe = pipe_roofline(dtype_name, stats, n_local, (device_ms - exch_ms) * 1e-3,
                         clock_info['sm_max_mhz'] if clock_info else None)
    roofline = {'bound': 'hbm', 'kernel': kernel_name, 'achieved': achieved, 'peak': peak,
                'unit': 'GB/s', 'frac': achieved / peak,
                'traffic': measured_traffic(dtype_name, n_local, kernel_name),
                'peak_source': peak_src, 'bytes_per_launch': pass_bytes,
                'avg_launch_ms': avg_pass_ms, 'launches_per_step': tile_passes / steps,
                'gates_per_launch': len(gates) * steps / max(1, tile_passes),
                'dense_ops_per_launch': (stats.get('shear_ops', 0) + stats.get('direct_ops', 0)) / max(1, tile_passes),
                # a pass that fuses many dense gates is bound by the FP pipe, not by HBM: both
                # fractions are reported, the larger one names the binding roofline
                'fp64_pipe_frac' if dtype_name == 'f64' else 'fp32_pipe_frac': pipe['frac'],
                'pipe': pipe}
    return {'value': upd * steps / (device_ms * 1e-3), 'device_ms': device_ms, 'stats': stats,
            'clocks': clock_info, 'roofline': roofline, 'p0': p0, 'exch_ms': exch_ms, 'dstats': dstats,
            'updates': upd}


def qft_records(n_gpus, g_bits, dtype_name, rank):
    """The other half of BASELINE.json's metric: QFT time.  At N GPUs: QFT-(30 + log2 N) (16 GiB per
    GPU, the bench shard) and the largest QFT that fits N GPUs up to 35 qubits (128 GiB per GPU at
    N = 1, 2, 4; 64 GiB at N = 8), each checked against the closed form of QFT|5> and P(0) = 1/2 of
    every qubit (run_configs.py)."""
    import run_configs
    import torch
    world = int(os.environ.get('WORLD_SIZE', '1'))
    records = []
    for n in sorted({30 + g_bits, min(33 + g_bits, 35)}):
        qargs = argparse.Namespace(qubits=n, dtype=dtype_name)
        try:
            # two runs, the faster one reported (both are in the record): the first pays the device
            # allocation of the state vector (and, sharded, the peer mappings) a warmed-up process re-uses
            first = run_configs.run_qft(qargs, torch, world, rank)
            second = run_configs.run_qft(qargs, torch, world, rank)
            rec = second if second['run_s'] <= first['run_s'] else first
            records.append({'qubits': n, 'gates': rec['gates'], 'state_bytes_per_gpu': rec['state_bytes'] // n_gpus,
                            'ms': 1e3 * rec['run_s'], 'first_run_ms': 1e3 * first['run_s'],
                            'second_run_ms': 1e3 * second['run_s'],
                            'all_qubit_p0_ms': 1e3 * rec['calc_probability_all_s'],
                            'amplitude_rel_err_vs_closed_form': rec['amplitude_rel_err_vs_closed_form'],
                            'p0_max_abs_err': rec['p0_max_abs_err'], 'ok': rec['ok']})
        except Exception as exc:   # e.g. not enough free device memory for the 128 GiB case
            records.append({'qubits': n, 'ok': False, 'error': str(exc)[:200]})
    return records


def main():
    args = parse_args()
    dtype = np.float64 if args.dtype == 'f64' else np.float32
    elem = 16 if args.dtype == 'f64' else 8
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    n_gpus = max(args.gpus, world)
    g_bits = int(round(math.log2(n_gpus)))
    if (1 << g_bits) != n_gpus:
        raise SystemExit('--gpus must be a power of two')
    n = args.qubits + g_bits            # weak scaling: fixed shard, more global qubits
    config = {'workload': 'random U3 + CX-ladder circuit (BASELINE configs[1] generator, seed {}) '
                          '{} qubits x depth {}, complex{}'.format(args.seed, n, args.depth,
                                                                   8 * elem),
              'qubits': n, 'depth': args.depth, 'state_bytes_per_gpu': elem << args.qubits,
              'sharding': 'global qubits, {} shard(s)'.format(n_gpus),
              'l2': 'state vector ({} GiB per GPU) far exceeds the 126 MB L2'.format(
                  (elem << args.qubits) >> 30)}

    if args.impl == 'reference':
        if rank != 0:
            return 0
        # the CPU arm runs the single-GPU problem size: one host has to hold the whole state and
        # sweep it once per gate.  Its line names what it ran.
        n_ref = min(n, args.qubits)
        if n_ref != n:
            config = dict(config)
            config['workload'] = config['workload'].replace('{} qubits'.format(n), '{} qubits'.format(n_ref))
            config['qubits'] = n_ref
            config['sharding'] = 'none (host memory)'
            config['note'] = ('the GPU arm at {} GPUs runs {} qubits (weak scaling, {} per GPU); the CPU arm '
                              'runs the 1-GPU size, {} qubits: updates/s is a rate, the sizes differ'
                              .format(n_gpus, n, args.qubits, n_ref))
        value, ms, desc = run_reference_arm(args, dtype, n_ref)
        line = {'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT,
                'n_gpus': n_gpus, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms,
                'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
                'dtype': args.dtype, 'data': 'synthetic', 'config': config,
                'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': cores_used(),
                                 'kind': 'reference', 'sample': desc},
                'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0,
                        'd2h_bytes_per_step': 0},
                'gpu_launches': 0}
        print(json.dumps(line))
        return 0

    import torch
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device: the engine has no CPU fallback')
    torch.cuda.set_device(local_rank)
    distributed = world > 1
    if distributed:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))

    import qgate_b200
    import qgate_b200.script as S
    from qgate_b200 import circuits, cudaruntime
    cudaruntime.set_preference(device_ids=[local_rank])
    api = cudaruntime.get_api()
    for opt in args.option:
        name, value = opt.split('=')
        api.set_option(name, int(value))

    gates = random_circuit_gates(n, args.depth, args.seed)
    upd = updates_of(gates, n)

    def barrier():
        torch.cuda.synchronize()
        if distributed:
            dist.barrier()
            torch.cuda.synchronize()

    # -- device-timed leg: native layer, gates queued first, events around the flush ------------
    if distributed:
        from qgate_b200 import dist as qdist
        runtime = qdist.runtime(cudaruntime, exchange='p2p' if args.exchange == 'p2p_inplace' else args.exchange,
                                push=args.exchange != 'p2p_inplace')
        runtime.ctx.bind_stream()              # NCCL work and the engine share this stream
        runtime.ctx.timing = True
        stream = runtime.ctx.stream
    else:
        dist = None
        runtime = cudaruntime
        cudaruntime.module_init() if not cudaruntime.initialized else None
        stream = torch.cuda.Stream()
        api.set_stream(stream.cuda_stream)
    leg = device_leg(args, args.dtype, n, gates, runtime, api, torch, dist, distributed, stream, args.steps,
                     args.warmup, True, local_rank)
    value, device_ms, stats, clock_info = leg['value'], leg['device_ms'], leg['stats'], leg['clocks']
    roofline, p0, exch_ms = leg['roofline'], leg['p0'], leg['exch_ms']
    achieved = roofline['achieved']
    launches = stats['kernel_launches']
    nvlink = None
    if distributed:
        dstats = leg['dstats']
        if dstats['exchanges']:
            # bytes each GPU sends (= receives) per exchange over its NVLink ports
            gbs = dstats['exchange_bytes'] / (exch_ms * 1e-3) / 1e9
            nvlink = {'exchange': runtime.ctx.exchange + (' (push into spare buffers)' if dstats.get('push_exchanges') else
                                                          ' (in place)' if runtime.ctx.exchange == 'p2p' else ''),
                      'exchanges_per_step': dstats['exchanges'] / args.steps,
                      'lanes_per_exchange': dstats['exchange_lanes'] / dstats['exchanges'],
                      'local_swaps_per_step': dstats['local_swaps'] / args.steps,
                      'bytes_per_gpu_per_direction_per_step': dstats['exchange_bytes'] // args.steps,
                      'ms_per_step': exch_ms / args.steps, 'achieved': gbs, 'unit': 'GB/s',
                      'peak': 770., 'frac': gbs / 770.,
                      'peak_source': 'measured peer copy per direction (B200_PROFILING.md)'}
    # -- the other precision of the same workload (sub-record; the headline stays args.dtype) ----
    other = None
    if not args.no_extras:
        other_dtype = 'f32' if args.dtype == 'f64' else 'f64'
        o_steps, o_warm = max(1, min(args.steps, 5)), max(1, min(args.warmup, 3))
        oleg = device_leg(args, other_dtype, n, gates, runtime, api, torch, dist, distributed, stream, o_steps,
                          o_warm, True, local_rank)
        other = {'dtype': other_dtype, 'value': oleg['value'], 'unit': UNIT, 'steps': o_steps, 'warmup': o_warm,
                 'ms_per_step': oleg['device_ms'] / o_steps, 'roofline': oleg['roofline'],
                 'clocks': oleg['clocks'], 'p0_check': oleg['p0'],
                 'gpu_launches': oleg['stats']['kernel_launches']}
    if not distributed:
        api.set_stream(0)

    # -- e2e leg: the call a qgate user makes, host objects in, NumPy out -------------------------
    # Front end: the REFERENCE's own (qgate.simulator.cuda(), unmodified, from baseline/_ref) with this
    # package's runtime installed as qgate.simulator.cudaruntime (qgate_b200/install.py) — its
    # per-gate Python dispatch is inside the timed region.  Without baseline/_ref: this package's
    # mirror of that front end.
    e2e = None
    if not args.no_e2e:
        from baseline import reference_frontend
        use_reference_front = reference_frontend.available() and not args.own_frontend
        if use_reference_front:
            import qgate_b200.install
            qgate = reference_frontend.load()
            front_runtime = runtime if distributed else cudaruntime
            qgate_b200.install.install(qgate, front_runtime)
            import qgate.script as RS
            q, ops = circuits.random_u3_cx(RS, n, args.depth, seed=args.seed)

            def make_sim():
                return qgate.simulator.cuda(dtype=dtype, circuit_prep=qgate.prefs.one_static)
            api_name = ('qgate.simulator.cuda().run(circuit) [reference front end, unmodified, over '
                        'qgate_b200.install]; qubits.calc_probability; qubits.states[:4096]')
        else:
            q, ops = circuits.random_u3_cx(S, n, args.depth, seed=args.seed)

            def make_sim():
                prefs = {'sharding': {'exchange': 'p2p' if args.exchange == 'p2p_inplace' else args.exchange,
                                      'push': args.exchange != 'p2p_inplace'}} if distributed else {}
                return qgate_b200.simulator.cuda(dtype=dtype, circuit_prep=qgate_b200.prefs.one_static, **prefs)
            api_name = 'qgate_b200.simulator.cuda().run(circuit); qubits.calc_probability; qubits.states[:4096]'
        e2e_steps = max(1, min(args.steps, 3))
        readback = 4096

        split = {'create_s': 0., 'run_s': 0., 'observe_s': 0., 'terminate_s': 0.}

        def e2e_step():
            t0 = time.perf_counter()
            sim = make_sim()
            t1 = time.perf_counter()
            sim.run(ops)                       # returns after the device has finished (synchronize)
            t2 = time.perf_counter()
            sim.qubits.set_ordering(q)
            p = sim.qubits.calc_probability(q[n - 1])
            head = sim.qubits.states[:readback]
            t3 = time.perf_counter()
            sim.terminate()
            del sim
            t4 = time.perf_counter()
            for key, dt in zip(('create_s', 'run_s', 'observe_s', 'terminate_s'),
                               (t1 - t0, t2 - t1, t3 - t2, t4 - t3)):
                split[key] += dt
            return p, head

        e2e_step()                        # warm-up (allocator, planner caches)
        import gc
        gc.collect()                      # the garbage of the legs above is not part of a user's run
        barrier()
        api.stats_reset()
        for key in split:
            split[key] = 0.
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            p, head = e2e_step()
        barrier()
        e2e_s = time.perf_counter() - t0
        if distributed:                          # slowest rank
            tt = torch.tensor([e2e_s], dtype=torch.float64, device='cuda')
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            e2e_s = float(tt.item())
        st = api.stats()
        assert abs(float(np.sum(np.abs(head) ** 2))) <= 1. and 0. <= p <= 1.
        e2e = {'value': upd * e2e_steps / e2e_s, 'unit': UNIT,
               'h2d_bytes_per_step': st['h2d_bytes'] // e2e_steps,
               'd2h_bytes_per_step': st['d2h_bytes'] // e2e_steps,
               'ms_per_step': 1e3 * e2e_s / e2e_steps, 'steps': e2e_steps,
               'split_ms': {k: 1e3 * v / e2e_steps for k, v in split.items()},
               'front_end': 'reference (baseline/_ref)' if use_reference_front else 'qgate_b200 mirror',
               'api': api_name, 'p0_check': p}

    # -- QFT sub-records (the metric's second half) -------------------------------------------------
    qft = None
    if not args.no_extras:
        qft = qft_records(n_gpus, g_bits, args.dtype, rank)

    # -- CPU baseline (rank 0, N=1 only) --------------------------------------------------------
    cpu_baseline = None
    if rank == 0 and not distributed and not args.no_cpu_baseline:
        try:
            sample = cpu_sample(gates, n, args.cpu_sample_gates)
            ref = ReferenceCpu(n, dtype)
            ref.apply(sample[:1])                   # touch the pages
            t0 = time.perf_counter()
            ref.apply(sample)
            dt = time.perf_counter() - t0
            ref.close()
            cpu_baseline = {'value': updates_of(sample, n) / dt, 'unit': UNIT,
                            'cores': cores_used(), 'kind': 'reference',
                            'sample': '{} gates of layer 0 (U3 on low/mid/high lanes + CX) at {} '
                                      'qubits, {:.1f} s'.format(len(sample), n, dt)}
        except Exception as exc:  # the checker is optional for the bench line, never for parity
            cpu_baseline = {'value': None, 'unit': UNIT, 'cores': cores_used(),
                            'kind': 'reference', 'sample': 'unavailable: {}'.format(exc)}

    if rank == 0:
        line = {'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': n_gpus,
                'steps': args.steps, 'warmup': args.warmup,
                'ms_per_step': device_ms / args.steps, 'higher_is_better': True,
                'scaling': 'weak', 'vs_baseline': None, 'dtype': args.dtype, 'data': 'synthetic',
                'config': config, 'clocks': clock_info, 'e2e': e2e,
                'gpu_launches': launches, 'roofline': roofline, 'cpu_baseline': cpu_baseline,
                'hbm_gbs_per_gpu': achieved, 'nvlink': nvlink,
                'gates': len(gates), 'p0_check': p0,
                # sub-records, not headlines: the other precision of the same workload, and the QFT
                # half of BASELINE.json's metric
                ('f32' if args.dtype == 'f64' else 'f64'): other, 'qft': qft}
        print(json.dumps(line))
    if distributed:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == '__main__':
    sys.exit(main())
