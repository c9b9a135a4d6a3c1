#!/usr/bin/env python
"""run_configs.py — BASELINE.json configs 3, 4 and 5 on the engine, with the checks that need no
512 GiB oracle.  One JSON line per config on rank 0.

    python run_configs.py grover [--qubits 30] [--iterations 8] [--shots 1000000]
    torchrun --nproc-per-node 2 run_configs.py pe [--qubits 32]           # 31 counting bits + target
    torchrun --nproc-per-node 8 run_configs.py qft [--qubits 35]

Single process = one GPU; under torchrun (one rank per GPU, NCCL) the state vector is sharded on
its high-order qubits (qgate_b200/dist.py).  Checks:

  grover  the marked state's probability against the closed form sin^2((2k+1) asin(2^-n/2)),
          the share of the 1M sampled indices that hit it, all-qubit measurement reproducing a
          basis state whose probability is consistent;
  pe      the histogram peak of the sampled counting register at round(v_in * 2^bits)
          (examples/phase_estimation.py:44-65) and its closed-form probability;
  qft     norm, every qubit's P(0) = 1/2, and amplitude slices against the closed form of
          QFT|x> validated against the reference CPU runtime at <= 20 qubits
          (tests/test_gpu_parity.py::test_large_state_properties).

Timing: wall clock around sim.run + the observers, device drained on both sides (reported as
info — bench.py holds the contract metric).
"""
import argparse
import json
import math
import os
import sys
import time

REPO = os.path.dirname(os.path.abspath(__file__))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

import numpy as np  # noqa: E402


def setup():
    import torch
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('run_configs.py needs a CUDA device: the engine has no CPU fallback')
    torch.cuda.set_device(local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    from qgate_b200 import cudaruntime
    cudaruntime.set_preference(device_ids=[local_rank])
    return torch, world, rank


def drain(torch, world):
    torch.cuda.synchronize()
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        torch.cuda.synchronize()


def run_grover(args, torch, world, rank):
    import qgate_b200
    import qgate_b200.script as S
    from qgate_b200 import circuits
    n, k = args.qubits, args.iterations
    dtype = np.float64 if args.dtype == 'f64' else np.float32
    marked = 0x2AAAAAAA & ((1 << n) - 1)
    q, ops = circuits.grover(S, n, k, marked)
    sim = qgate_b200.simulator.cuda(dtype=dtype, circuit_prep=qgate_b200.prefs.one_static)
    drain(torch, world)
    import gc
    gc.collect()           # (a full collection of the caller's garbage inside a 50 ms run is a 100 ms outlier)
    gc.disable()
    try:
        t0 = time.perf_counter()
        sim.run(ops)
        sim.qubits.set_ordering(q)
        drain(torch, world)
        t_run = time.perf_counter() - t0
    finally:
        gc.enable()
    theta = math.asin(2. ** (-n / 2.))
    p_want = math.sin((2 * k + 1) * theta) ** 2
    p_got = float(sim.qubits.prob[marked])
    t0 = time.perf_counter()
    pool = sim.qubits.create_sampling_pool(q)
    drain(torch, world)
    t_pool = time.perf_counter() - t0
    rnd = np.random.RandomState(7).random_sample(args.shots)
    t0 = time.perf_counter()
    obs = pool.sample(args.shots, rnd).intarray
    drain(torch, world)
    t_sample = time.perf_counter() - t0
    hits = int(np.count_nonzero(obs == marked))
    sigma = math.sqrt(args.shots * p_want * (1. - p_want)) + 1.
    # measure every qubit (np.random.seed(11): one draw per Measure)
    refs = S.new_references(n)
    np.random.seed(11)
    t0 = time.perf_counter()
    sim.run([S.measure(r, qr) for r, qr in zip(refs, q)])
    drain(torch, world)
    t_measure = time.perf_counter() - t0
    bits = sim.values.get(refs)
    outcome = sum(int(b) << i for i, b in enumerate(bits))
    collapsed = sim.qubits.states[outcome]
    ok = (abs(p_got - p_want) < (1e-12 if dtype is np.float64 else 1e-6)
          and abs(hits - args.shots * p_want) < 6. * sigma
          and abs(abs(collapsed) - 1.) < (1e-9 if dtype is np.float64 else 1e-4)
          and int(obs.min()) >= 0 and int(obs.max()) < (1 << n))
    sim.terminate()
    return {'config': 'grover', 'qubits': n, 'iterations': k, 'dtype': args.dtype, 'gates': len(ops),
            'marked': marked, 'p_marked': p_got, 'p_marked_closed_form': p_want,
            'shots': args.shots, 'hits_marked': hits, 'hits_expected': args.shots * p_want,
            'measured_outcome': outcome, 'collapsed_amplitude_abs': float(abs(collapsed)),
            'run_s': t_run, 'pool_build_s': t_pool, 'sample_s': t_sample, 'measure_all_s': t_measure,
            'ok': bool(ok)}


def run_pe(args, torch, world, rank):
    import qgate_b200
    import qgate_b200.script as S
    from qgate_b200 import circuits
    n_bits = args.qubits - 1
    v_in = 0.1
    dtype = np.float64 if args.dtype == 'f64' else np.float32
    bits, target, ops = circuits.phase_estimation(S, n_bits, v_in)
    sim = qgate_b200.simulator.cuda(dtype=dtype, circuit_prep=qgate_b200.prefs.one_static)
    drain(torch, world)
    t0 = time.perf_counter()
    sim.run(ops)
    drain(torch, world)
    t_run = time.perf_counter() - t0
    t0 = time.perf_counter()
    pool = sim.qubits.create_sampling_pool(bits)       # the target is a hidden lane
    rnd = np.random.RandomState(3).random_sample(args.shots)
    obs = pool.sample(args.shots, rnd).intarray
    drain(torch, world)
    t_sample = time.perf_counter() - t0
    values, counts = np.unique(obs, return_counts=True)
    peak = int(values[np.argmax(counts)])
    size = float(1 << n_bits)
    # the inverse QFT has no final swaps: counting bit idx weighs 2^-(idx+1), like to_real() of
    # examples/phase_estimation.py:36-42, so the register reads the estimate bit-reversed
    natural = int(round(v_in * size)) % (1 << n_bits)
    want = int(format(natural, '0{}b'.format(n_bits))[::-1], 2)
    # closed form: the counting register ends in sum_k |k> (1/2^n) sum_j e^{2 pi i j (v - k/2^n)}
    delta = v_in - natural / size
    p_peak = (math.sin(math.pi * size * delta) / (size * math.sin(math.pi * delta))) ** 2 \
        if delta != 0. else 1.
    share = float(counts.max()) / args.shots
    sigma = math.sqrt(p_peak * (1. - p_peak) / args.shots) + 1. / args.shots
    ok = peak == want and abs(share - p_peak) < 6. * sigma
    sim.terminate()
    return {'config': 'phase_estimation', 'qubits': args.qubits, 'counting_bits': n_bits,
            'dtype': args.dtype, 'gates': len(ops), 'v_in': v_in, 'peak': peak, 'peak_expected': want,
            'estimate': sum(((peak >> idx) & 1) * 0.5 ** (idx + 1) for idx in range(n_bits)), 'peak_share': share, 'peak_probability_closed_form': p_peak,
            'shots': args.shots, 'run_s': t_run, 'pool_and_sample_s': t_sample, 'ok': bool(ok)}


def run_qft(args, torch, world, rank):
    import qgate_b200
    import qgate_b200.script as S
    from qgate_b200 import circuits
    n = args.qubits
    dtype = np.float64 if args.dtype == 'f64' else np.float32
    q, ops = circuits.qft(S, n)                 # X(q0) X(q2) -> |5>, then the QFT gate list
    sim = qgate_b200.simulator.cuda(dtype=dtype, circuit_prep=qgate_b200.prefs.one_static)
    drain(torch, world)
    t0 = time.perf_counter()
    sim.run(ops)
    sim.qubits.set_ordering(q)
    drain(torch, world)
    t_run = time.perf_counter() - t0
    amp = 2. ** (-n / 2.)
    tol = 1e-12 if dtype is np.float64 else 1e-5
    t0 = time.perf_counter()
    p0 = [sim.qubits.calc_probability(qr) for qr in q]
    drain(torch, world)
    t_prob = time.perf_counter() - t0
    # amplitude of |k>: exp(2 pi i k rev_n(x) / 2^n) / sqrt(2^n), x = 5 (no final swaps)
    step = (1 << max(0, n - 16)) + 1
    start = 12345 % (1 << n)
    got = sim.qubits.states[start::step]
    ks = np.arange(start, 1 << n, step, dtype=np.int64)
    xr = (1 << (n - 1)) + (1 << (n - 3))
    phase = ((ks % (1 << n)) * (xr % (1 << n))) % (1 << n) if n <= 31 else \
        np.array([(int(k) * xr) % (1 << n) for k in ks], dtype=np.float64)
    want = amp * np.exp(2j * np.pi * np.asarray(phase, np.float64) / float(1 << n))
    err = float(np.abs(got - want).max() / amp)
    p0_err = float(np.abs(np.array(p0) - 0.5).max())
    ok = err < tol and p0_err < tol
    sim.terminate()
    return {'config': 'qft', 'qubits': n, 'dtype': args.dtype, 'gates': len(ops),
            'state_bytes': (16 if dtype is np.float64 else 8) << n, 'run_s': t_run,
            'calc_probability_all_s': t_prob, 'slice_points': int(ks.size),
            'amplitude_rel_err_vs_closed_form': err, 'p0_max_abs_err': p0_err, 'ok': bool(ok)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('config', choices=('grover', 'pe', 'qft'))
    ap.add_argument('--qubits', type=int, default=None)
    ap.add_argument('--iterations', type=int, default=8)
    ap.add_argument('--shots', type=int, default=1000000)
    ap.add_argument('--dtype', default='f64', choices=('f64', 'f32'))
    ap.add_argument('--option', action='append', default=[], help='engine option name=value')
    ap.add_argument('--repeat', type=int, default=1, help='run the config this many times, report the last '
                    '(the first run pays the device allocation of the state vector)')
    args = ap.parse_args()
    if args.qubits is None:
        args.qubits = {'grover': 30, 'pe': 32, 'qft': 35}[args.config]
    torch, world, rank = setup()
    from qgate_b200 import cudaruntime
    api = cudaruntime.get_api()
    for opt in args.option:
        name, value = opt.split('=')
        api.set_option(name, int(value))
    for _ in range(max(1, args.repeat)):
        api.stats_reset()
        out = {'grover': run_grover, 'pe': run_pe, 'qft': run_qft}[args.config](args, torch, world, rank)
    out['n_gpus'] = world
    out['engine_stats'] = api.stats()
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()
    return 0 if out['ok'] else 1


if __name__ == '__main__':
    sys.exit(main())
