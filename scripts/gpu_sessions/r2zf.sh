#!/bin/bash
# round-2 GPU session zf (1 GPU, last): the shared-GPU sharded suite with the kept IPC mappings, a subset of the
# parity suite, smoke, a short bench line
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_dist_shared_gpu.py -m gpu -q -x ) > gpurun_out/r2zf_pytest_shared.log 2>&1; tail -3 gpurun_out/r2zf_pytest_shared.log
( timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_qasm_reader.py tests/test_native_multi_qubit.py -m gpu -q -x -k "not 24_qubits and not 28_qubits and not grover_22 and not phase_estimation_20" ) > gpurun_out/r2zf_pytest_parity.log 2>&1; tail -2 gpurun_out/r2zf_pytest_parity.log
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 300 python bench.py --steps 3 --warmup 3 --depth 60 --no-cpu-baseline > gpurun_out/r2zf_bench_d60.json 2> gpurun_out/r2zf_bench_d60.err; python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/r2zf_bench_d60.json').read().strip().splitlines()[-1])
    r = d['roofline']
    print('f64 upd/s %.3e frac %.3f | e2e %.3e | f32 %.3e | qft %s' % (d['value'], r['frac'], d['e2e']['value'], d['f32']['value'], [(q['qubits'], round(q.get('ms', -1), 1), q['ok']) for q in d['qft']]))
except Exception as e:
    print('bench failed', e, open('gpurun_out/r2zf_bench_d60.err').read()[-600:])
PY
