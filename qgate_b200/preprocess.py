"""Circuit preprocessing: macro expansion and dynamic qubit grouping.

The native layer only ever sees single-qubit (multi-)controlled gates, single-lane
prob / measure, join, separate and reset.  This module produces that stream:

 * Swap -> 3 CX; Expi(theta)(Pauli string) and Pauli measure / prob -> basis change +
   CX chain + a single-qubit op + undo   (qgate/model/expand.py:14-90,
   qgate/model/pauli_gates_diagonalizer.py:10-146)
 * grouping directives NewQreg / Join / Separate according to the `circuit_prep`
   preference                            (qgate/model/preprocessor.py:8-142,
                                          qgate/model/qreg_aggregator.py:2-71)
"""
from . import model
from .model import gate_type as gtype
from .model import prefs


# ---- Pauli-string diagonalisation ------------------------------------------------------
def _gate(gate_type, qreg, ctrl=None):
    gate = model.Gate(gate_type)
    if ctrl is not None:
        gate.set_ctrllist([ctrl])
    gate.set_qreg(qreg)
    return gate


# Reduced X/Z word on one lane (adjacent equal letters cancel, no commutation) ->
# (extra phase in units of pi/2, basis-change gate or None).  The word is the product of
# the lane's Pauli gates with Y written as 'XZ' (+1 phase unit); '(ZX)^2' = -1 folds the
# word length mod 8.  Empty word = identity.
_WORD_TABLE = {
    'Z': (0, None), 'X': (0, gtype.H),
    'ZX': (1, gtype.SH), 'XZ': (3, gtype.SH),
    'ZXZ': (2, gtype.H), 'XZX': (2, None),
}


def _reduce_lane(gates):
    phase, word = 0, []
    for gate in gates:
        gt = gate.gate_type
        if isinstance(gt, gtype.X):
            letters = 'X'
        elif isinstance(gt, gtype.Y):
            letters, phase = 'XZ', phase + 1
        elif isinstance(gt, gtype.Z):
            letters = 'Z'
        elif isinstance(gt, gtype.ID):
            letters = ''
        else:
            raise RuntimeError('input must be X, Y, Z or ID, but {} passed.'.format(gt))
        for letter in letters:
            if word and word[-1] == letter:
                word.pop()
            else:
                word.append(letter)
    pairs, trailing = divmod(len(word), 2)
    pairs %= 4
    if pairs >= 2:
        phase, pairs = phase + 2, pairs - 2
    core = ''.join(word[:2 * pairs + trailing])
    if not core:
        return phase, False, None
    extra, basis = _WORD_TABLE[core]
    return phase + extra, True, basis


class PauliDiagonalizer:
    """Maps a Pauli string onto +-(i) Z (or I) on one qreg by basis changes + a CX chain."""

    def __init__(self, gatelist):
        per_qreg = dict()
        for gate in gatelist:
            per_qreg.setdefault(gate.qreg, []).append(gate)
        phase, zlanes, idlanes, basis_gates = 0, [], [], []
        for qreg in sorted(per_qreg, key=lambda q: q.id):
            lane_phase, z_based, basis = _reduce_lane(per_qreg[qreg])
            phase += lane_phase
            if z_based:
                zlanes.append(qreg)
                if basis is not None:
                    basis_gates.append(_gate(basis(), qreg))
            else:
                idlanes.append(qreg)
        self.phase_units = phase % 4
        self.z_based = len(zlanes) != 0
        if self.z_based:
            chain = [_gate(gtype.X(), zlanes[i + 1], zlanes[i]) for i in range(len(zlanes) - 1)]
            self.op_qreg = zlanes[-1]
        else:
            chain = []
            self.op_qreg = idlanes[0]
        self.pcx = basis_gates + chain

    def phase_coef(self):
        return (1, 1j, -1, -1j)[self.phase_units]


def reduce_pauli_string(gatelist):
    """Pauli gate list (several gates per qreg allowed) -> (coefficient in {1, 1j, -1, -1j},
    [(qreg, code)] with code 0 = I, 1 = X, 2 = Y, 3 = Z, one entry per qreg, ascending qreg id):
    the product of the list equals coefficient x the tensor product of the coded Paulis.  Same
    per-lane reduction as PauliDiagonalizer (pauli_gates_diagonalizer.py:96-146): a lane whose word
    reduces to Z / X / +-Y is diagonalised by nothing / H / SH there."""
    per_qreg = dict()
    for gate in gatelist:
        per_qreg.setdefault(gate.qreg, []).append(gate)
    phase, out = 0, []
    for qreg in sorted(per_qreg, key=lambda q: q.id):
        lane_phase, z_based, basis = _reduce_lane(per_qreg[qreg])
        phase += lane_phase
        code = 0 if not z_based else 3 if basis is None else 1 if basis is gtype.H else 2
        out.append((qreg, code))
    return (1, 1j, -1, -1j)[phase % 4], out


def _adjoint(gates):
    out = []
    for gate in reversed(gates):
        dup = gate.copy()
        dup.adjoint = not gate.adjoint
        out.append(dup)
    return out


def _sandwich(diag, middle):
    seq = [g.copy() for g in diag.pcx] + [middle] + _adjoint(diag.pcx)
    seq.reverse()
    return seq


def expand_swap(op):
    q0, q1 = op.qreglist
    return [_gate(gtype.X(), q1, q0), _gate(gtype.X(), q0, q1), _gate(gtype.X(), q1, q0)]


def expand_exp(exp):
    diag = PauliDiagonalizer(exp.gatelist)
    if diag.phase_units % 2 != 0:
        raise RuntimeError('cannot expand, {}.'.format(repr(exp)))
    theta = (diag.phase_coef() * exp.gate_type.args[0]).real
    if exp.adjoint:
        # exp(i theta P)^+ = exp(-i theta P).  (The reference raises NameError here,
        # qgate/model/expand.py:45-47.)
        theta = -theta
    middle = _gate((gtype.ExpiZ if diag.z_based else gtype.ExpiI)(theta), diag.op_qreg)
    seq = _sandwich(diag, middle)
    if exp.ctrllist is not None:
        ctrlset = set(exp.ctrllist)
        for gate in seq:
            merged = ctrlset | set(gate.ctrllist or [])
            assert gate.qreg not in merged, 'control bits must not overlap targets.'
            gate.set_ctrllist(list(merged))
    return seq


def expand_pauli_observer(op):
    diag = PauliDiagonalizer(op.gatelist)
    if not diag.z_based:
        raise RuntimeError('measurement is not z-based.')
    cls = model.Measure if isinstance(op, model.PauliMeasure) else model.Prob
    return _sandwich(diag, cls(op.outref, diag.op_qreg))


def expand(op):
    if isinstance(op, (model.PauliMeasure, model.PauliProb)):
        return expand_pauli_observer(op)
    if isinstance(op.gate_type, gtype.SWAP):
        return expand_swap(op)
    if isinstance(op.gate_type, gtype.Expi):
        return expand_exp(op)
    raise RuntimeError('Unknown composed gate, {}.'.format(repr(op.gate_type)))


# ---- grouping ----------------------------------------------------------------------------
class QregGroups:
    """Tracks which qregs currently share a state vector."""

    def __init__(self):
        self.groups = []          # list of frozensets
        self.qregset = set()

    def find(self, qreg):
        for group in self.groups:
            if qreg in group:
                return group
        raise RuntimeError('qreg{} is not in circuit.'.format(qreg.id))

    def add(self, qreg):
        if qreg in self.qregset:
            return False
        self.qregset.add(qreg)
        self.groups.append(frozenset([qreg]))
        return True

    def merge(self, qregs):
        for qreg in qregs:
            self.add(qreg)
        touched = set(self.find(qreg) for qreg in qregs)
        if len(touched) == 1:
            return None
        merged = set()
        for group in touched:
            self.groups.remove(group)
            merged |= group
        self.groups.append(frozenset(merged))
        return merged

    def separate(self, qreg):
        group = self.find(qreg)
        self.groups.remove(group)
        self.groups.append(frozenset(group - {qreg}))
        self.groups.append(frozenset([qreg]))


class Preprocessor:
    def __init__(self, **prefdict):
        self.circ_prep = prefdict.get(prefs.circuit_prep, prefs.dynamic)
        if self.circ_prep not in (prefs.dynamic, prefs.static, prefs.one_static):
            raise RuntimeError('unknown circuit_prep, {}.'.format(self.circ_prep))
        self.dynamic = self.circ_prep == prefs.dynamic
        # the runtime applies Swap and exp(i theta P) natively: they stay
        # whole, only the grouping they imply is recorded
        self.native_multi = bool(prefdict.get('native_multi_qubit_ops', False))
        self.reset()

    def reset(self):
        self.groups = QregGroups()
        self._refset = set()

    def get_refset(self):
        return self._refset

    def get_qregset(self):
        return self.groups.qregset

    def _group(self, qregs, out):
        """the qregs of one multi-qubit op must share a state vector"""
        if len(qregs) == 1:
            if self.groups.add(qregs[0]) and self.dynamic:
                out.append(model.NewQreg(qregs[0]))
            return
        merged = self.groups.merge(qregs)
        if merged is not None and self.dynamic:
            out.append(model.Join(merged))

    def _one(self, op, out):
        groups = self.groups
        if isinstance(op, model.Gate):
            if op.ctrllist is None:
                if groups.add(op.qreg) and self.dynamic:
                    out.append(model.NewQreg(op.qreg))
            else:
                merged = groups.merge([op.qreg] + op.ctrllist)
                if merged is not None and self.dynamic:
                    out.append(model.Join(merged))
            out.append(op)
        elif isinstance(op, (model.Measure, model.Prob)):
            if groups.add(op.qreg) and self.dynamic:
                out.append(model.NewQreg(op.qreg))
            self._refset.add(op.outref)
            out.append(op)
            if isinstance(op, model.Measure) and self.dynamic:
                groups.separate(op.qreg)
                out.append(model.Separate(op.qreg))
        elif isinstance(op, model.ReleaseQreg):
            if self.dynamic:
                if op.qreg not in groups.qregset:
                    raise RuntimeError('qreg{} is not in circuit.'.format(op.qreg.id))
                if len(groups.find(op.qreg)) != 1:
                    raise RuntimeError('qreg{} is not separated.'.format(op.qreg.id))
                out.append(op)
        elif isinstance(op, model.Reset):
            if op.qreg not in groups.qregset:
                raise RuntimeError('unused qreg found, {}'.format(op.qreg))
            out.append(op)
        elif isinstance(op, model.Barrier):
            out.append(op)
        elif self.native_multi and isinstance(op, model.MultiQubitGate) and \
                isinstance(op.gate_type, gtype.SWAP):
            self._group(list(op.qreglist), out)
            out.append(op)
        elif self.native_multi and isinstance(op, model.GatelistMacro) and \
                isinstance(op.gate_type, gtype.Expi):
            coef, _ = reduce_pauli_string(op.gatelist)
            if coef.imag != 0:
                raise RuntimeError('cannot expand, {}.'.format(repr(op)))
            qregs = []
            for gate in op.gatelist:
                if gate.qreg not in qregs:
                    qregs.append(gate.qreg)
            self._group(qregs + list(op.ctrllist or []), out)
            out.append(op)
        elif isinstance(op, (model.MultiQubitGate, model.GatelistMacro,
                             model.PauliMeasure, model.PauliProb)):
            for inner in expand(op):
                self._one(inner, out)
        elif isinstance(op, (model.GateList, list, tuple)):
            for inner in model.flatten(op):
                self._one(inner, out)
        elif isinstance(op, model.IfClause):
            body = []
            for inner in model.flatten(op.clause):
                self._one(inner, body)
            out.append(model.IfClause(op.refs, op.cond, body))
        else:
            raise RuntimeError('Unexpected op, {}'.format(repr(op)))

    def preprocess(self, circuit):
        """Returns a flat python list of operators (IfClause bodies are flat lists too)."""
        out = []
        self._one(circuit, out)
        if not self.dynamic:
            prologue = []
            if self.circ_prep == prefs.static:
                partitions = self.groups.groups
            else:
                partitions = [self.groups.qregset] if self.groups.qregset else []
            for group in partitions:
                if len(group) == 1:
                    prologue.append(model.NewQreg(*group))
                else:
                    prologue.append(model.Join(group))
            out = prologue + out
        return out
