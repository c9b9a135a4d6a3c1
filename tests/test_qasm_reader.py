"""the OpenQASM 2.0 front end (qgate_b200/openqasm.py) — the statement-level
cases of the reference's own tests (tests/test_openqasm.py: declarations, U / CX, qelib1 gates with
indexed / whole-register / mixed arguments, measure, reset, barrier, and its error behaviour), plus
what the reference's tests do not check: that the circuit it builds IS the program (amplitudes and
measured values against the same circuit written with the script API, on the reference CPU runtime)
and that translate() emits source that builds the same circuit."""
import math

import numpy as np
import pytest

import qgate_b200
import qgate_b200.script as S
from qgate_b200 import model, openqasm
from tests import cases



def test_empty_and_declarations():
    mod = openqasm.load_circuit('')
    assert mod.circuit == []
    mod = openqasm.load_circuit('OPENQASM 2.0;\ninclude "qelib1.inc";\nqreg q[10]; creg c[10];')
    assert len(mod.q) == 10 and len(mod.c) == 10 and mod.circuit == []
    assert all(isinstance(x, model.Qreg) for x in mod.q) and all(isinstance(x, model.Reference) for x in mod.c)


@pytest.mark.parametrize('qasm,count', [
    ('qreg q[10]; creg c[10]; U(0., 0., 0.) q[2]; U(0., 0., 0.) q;', 11),
    ('qreg q[10]; CX q[2], q[3];', 1),
    ('qreg q[10]; h q[2], q[3];', 2),
    ('qreg q[10]; h q[2], q;', 11),
    ('qreg q[10]; u1(0.) q[0]; u2(0., 0.) q[0]; u3(0., 0., 0.) q[0];', 3),
    ('qreg q[10]; creg c[10]; measure q[0] -> c[0];', 1),
    ('qreg q[10]; creg c[10]; measure q -> c;', 10),
    ('qreg q[10]; reset q[0];', 1),
    ('qreg q[10]; reset q;', 10),
    ('qreg q[10]; barrier q[0];', 1),
    ('qreg q[10]; barrier q;', 10),
    ('qreg q[3]; qreg r[3]; cx q, r; cx q[0], r; swap q[1], r[2]; ccx q[0], q[1], r[0];', 8),
    ('qreg q[2]; creg c[2]; measure q[0] -> c[0]; if (c == 1) x q[1];', 2),
    ('qreg q[1]; // a comment\nrx(pi / 2) q[0]; ry(-pi) q[0]; rz(2 * pi / 3) q[0]; sdg q[0]; tdg q[0]; id q[0];', 6),
])
def test_statements_of_the_reference_tests(qasm, count):
    mod = openqasm.load_circuit(qasm)
    assert len(mod.circuit) == count
    # the emitted source builds a circuit of the same length
    ns = {}
    exec(openqasm.translate(qasm), ns)
    flat = []

    def walk(items):
        for item in items:
            if isinstance(item, (list, tuple)):
                walk(item)
            else:
                flat.append(item)
    walk(ns.get('circuit', []))
    assert len(flat) == count


def test_error_behaviour_matches_the_reference():
    with pytest.raises(RuntimeError):        # tests/test_openqasm.py:21-25
        openqasm.load_circuit('qreg q[10]; creg c[10]; gate custome_gate() a, b, c { }')
    with pytest.raises(NotImplementedError):  # :27-31
        openqasm.load_circuit('qreg q[10]; creg c[10];\nopaque opaque_gate() a, b, c;')
    with pytest.raises(NameError):           # :110-115
        openqasm.load_circuit('qreg q[10]; barrier g;')
    with pytest.raises(SyntaxError) as exc:  # :117-126 (lexer)
        openqasm.load_circuit('qreg q[10];\nh q[0];\nh q[1];\n_h q[2];\nh q[3];\n')
    assert 'line 4' in str(exc.value)
    with pytest.raises(SyntaxError) as exc:  # :128-137 (parser)
        openqasm.load_circuit('qreg q[10];\nh q[0]\nh q[1];\n')
    assert 'line 3' in str(exc.value)
    with pytest.raises(RuntimeError):
        openqasm.load_circuit('qreg q[2]; nosuchgate q[0];')
    with pytest.raises(SyntaxError):
        openqasm.load_circuit('qreg q[2]; u3(1, 2) q[0];')
    with pytest.raises(IndexError):
        openqasm.load_circuit('qreg q[2]; h q[2];')


def test_expressions():
    mod = openqasm.load_circuit('qreg q[1]; u3(pi/2, -pi/4 + 1e-1, 2^3 * sin(0.5)) q[0]; u1(sqrt(2) - ln(exp(1.5))) q[0];')
    a, b = mod.circuit
    assert np.allclose(a.gate_type.args, (math.pi / 2, -math.pi / 4 + 0.1, 8 * math.sin(0.5)))
    assert np.allclose(b.gate_type.args, (math.sqrt(2.) - 1.5,))


def qft_qasm(n):
    """the QFT of examples/qft.qasm (no final swaps) generalised to n qubits, bit order reversed like
    the example: h q[n-1]; cu1(pi/2) q[n-2], q[n-1]; ..."""
    lines = ['OPENQASM 2.0;', 'include "qelib1.inc";', 'qreg q[{}];'.format(n), 'x q[0];', 'x q[2];']
    for t in range(n - 1, -1, -1):
        lines.append('h q[{}];'.format(t))
        for c in range(t - 1, -1, -1):
            lines.append('cu1(pi/{}) q[{}], q[{}];'.format(1 << (t - c), c, t))
    return '\n'.join(lines)


def qft_script(n):
    q = S.new_qregs(n)
    ops = [S.X(q[0]), S.X(q[2])]
    for t in range(n - 1, -1, -1):
        ops.append(S.H(q[t]))
        for c in range(t - 1, -1, -1):
            ops.append(S.ctrl(q[c]).U1(math.pi / float(1 << (t - c)))(q[t]))
    return q, ops


def run_states(runtime, q, ops, refs=None, seed=5):
    sim = cases.make_sim(runtime, np.float64, 'one_static')
    np.random.seed(seed)
    sim.run(ops)
    sim.qubits.set_ordering(q)
    out = sim.qubits.states[:], (sim.values.get(refs) if refs else None)
    sim.terminate()
    return out


def test_qft_program_equals_the_script_circuit(ref_runtime):
    n = 9
    mod = openqasm.load_circuit(qft_qasm(n))
    got, _ = run_states(ref_runtime.module, mod.q, mod.circuit)
    q, ops = qft_script(n)
    want, _ = run_states(ref_runtime.module, q, ops)
    assert np.abs(got - want).max() < 1e-14
    # and through translate() + exec
    ns = {}
    exec(openqasm.translate(qft_qasm(n)), ns)
    got2, _ = run_states(ref_runtime.module, ns['q'], ns['circuit'])
    assert np.abs(got2 - want).max() < 1e-14


def test_program_with_measurement_reset_if_and_broadcasts(ref_runtime):
    qasm = '''
    OPENQASM 2.0;
    include "qelib1.inc";
    qreg a[3]; qreg b[3]; creg m[3]; creg f[1];
    h a;
    cx a, b;                 // pairwise
    u3(0.3, 0.2, 0.1) b[1];
    rz(0.7) a[2];
    ccx a[0], b[0], b[2];
    cz a[1], b;              // one control against a whole register
    swap a[0], b[1];
    barrier a, b[0];
    measure a -> m;
    if (m == 5) y b[0];
    reset a[1];
    ch b[2], a[1];
    tdg b; sdg a[0];
    cu3(0.4, 0.5, 0.6) b[0], a[2];
    measure b[2] -> f[0];
    '''
    mod = openqasm.load_circuit(qasm)
    a, b = S.new_qregs(3), S.new_qregs(3)
    m, f = S.new_references(3), S.new_references(1)
    ops = [S.H(x) for x in a] + [S.ctrl(x).X(y) for x, y in zip(a, b)]
    ops += [S.U3(0.3, 0.2, 0.1)(b[1]), S.U1(0.7)(a[2]), S.ctrl(a[0], b[0]).X(b[2])]
    ops += [S.ctrl(a[1]).Z(y) for y in b]
    ops += [S.Swap(a[0], b[1])] + [S.barrier(x) for x in a] + [S.barrier(b[0])]
    ops += [S.measure(r, x) for r, x in zip(m, a)]
    ops += [S.if_(m, 5, [S.Y(b[0])]), S.reset(a[1]), S.ctrl(b[2]).H(a[1])]
    ops += [S.T.Adj(y) for y in b] + [S.S.Adj(a[0])]
    ops += [S.ctrl(b[0]).U3(0.4, 0.5, 0.6)(a[2]), S.measure(f[0], b[2])]
    for seed in (1, 2, 3):
        got, got_bits = run_states(ref_runtime.module, mod.a + mod.b, mod.circuit, mod.m + mod.f, seed)
        want, want_bits = run_states(ref_runtime.module, a + b, ops, m + f, seed)
        assert got_bits == want_bits
        assert np.abs(got - want).max() < 1e-14


@pytest.mark.gpu
def test_qasm_qft_on_the_engine(cuda_runtime, ref_runtime):
    """the same text on the CUDA engine (phase fans in the textbook gate order) and on the oracle"""
    n = 18
    mod = openqasm.load_circuit(qft_qasm(n))
    got, _ = run_states(cuda_runtime, mod.q, mod.circuit)
    mod2 = openqasm.load_circuit(qft_qasm(n))
    want, _ = run_states(ref_runtime.module, mod2.q, mod2.circuit)
    assert cases.rel_err(got, want) < 1e-12
