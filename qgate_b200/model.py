"""Circuit model: qregs, references, gate types, operators and directives.

This is the host-side vocabulary the script API builds and the simulator consumes.
It mirrors the public names of the reference's `qgate.model` package
(qgate/model/model.py, gate_type.py, directive.py, gatelist.py, prefs.py) so user
scripts port by changing one import; the implementation here is independent and
kept to what the state-vector hot path needs.
"""
import itertools


# --- preferences (qgate/model/prefs.py:1-10) -----------------------------------------
class prefs:
    circuit_prep = 'circuit_prep'
    dynamic = 'dynamic'        # groups are joined / separated as the circuit runs
    static = 'static'          # one state vector per connected component, decided up front
    one_static = 'one_static'  # a single state vector holding every qreg


# --- handles ---------------------------------------------------------------------------
class _Handle:
    """Integer-identified handle; hash == id like the reference (model.py:4-37), which
    makes set iteration order (and therefore lane assignment in joins) reproducible."""
    __slots__ = ('id',)
    _counter = None

    def __init__(self):
        self.id = next(type(self)._counter)

    def __hash__(self):
        return self.id

    def __eq__(self, other):
        return type(other) is type(self) and other.id == self.id

    def __ne__(self, other):
        return not self.__eq__(other)

    def __repr__(self):
        return '{}({})'.format(type(self).__name__, self.id)


class Qreg(_Handle):
    """Logical qubit."""
    __slots__ = ()
    _counter = itertools.count()


class Reference(_Handle):
    """Classical result slot (measured bit or probability)."""
    __slots__ = ()
    _counter = itertools.count()


# --- gate types (qgate/model/gate_type.py:35-130) --------------------------------------
class GateType:
    n_params = 0

    def __init__(self, *args):
        if len(args) != self.n_params:
            raise RuntimeError('{} takes {} parameter(s).'.format(type(self).__name__,
                                                                 self.n_params))
        self.args = tuple(args)

    def __repr__(self):
        if not self.args:
            return type(self).__name__
        return '{}({})'.format(type(self).__name__, ', '.join(repr(a) for a in self.args))


def _gate_type(name, n_params=0):
    return type(name, (GateType,), {'n_params': n_params})


class gate_type:
    """Namespace of the 16 single-qubit gate types plus the two composed ones."""
    GateType = GateType
    U = _gate_type('U', 3)
    U2 = _gate_type('U2', 2)
    U1 = _gate_type('U1', 1)
    ID = _gate_type('ID')
    X = _gate_type('X')
    Y = _gate_type('Y')
    Z = _gate_type('Z')
    H = _gate_type('H')
    S = _gate_type('S')
    T = _gate_type('T')
    RX = _gate_type('RX', 1)
    RY = _gate_type('RY', 1)
    RZ = _gate_type('RZ', 1)
    SH = _gate_type('SH')          # basis change X -> Z used by Pauli measurements
    ExpiI = _gate_type('ExpiI', 1)
    ExpiZ = _gate_type('ExpiZ', 1)
    Expi = _gate_type('Expi', 1)   # exp(i theta P), P a Pauli string; expanded before execution
    SWAP = _gate_type('SWAP')      # expanded to 3 CX

    PAULI = None                   # filled below


gate_type.PAULI = (gate_type.ID, gate_type.X, gate_type.Y, gate_type.Z)


# --- operators -------------------------------------------------------------------------
class Operator:
    def copy(self):
        return self


def _as_qreg_list(qregs):
    qregs = [qregs] if isinstance(qregs, Qreg) else list(qregs)
    for qreg in qregs:
        if not isinstance(qreg, Qreg):
            raise RuntimeError('{} is not a qreg'.format(repr(qreg)))
    return qregs


class Gate(Operator):
    """Single-qubit gate with optional control qregs (model.py:55-88)."""

    def __init__(self, gate_type):
        self.gate_type = gate_type
        self.adjoint = False
        self.qreg = None
        self.ctrllist = None

    def set_adjoint(self, adjoint):
        self.adjoint = adjoint

    def set_ctrllist(self, ctrllist):
        self.ctrllist = _as_qreg_list(ctrllist)

    def set_qreg(self, qreg):
        if not isinstance(qreg, Qreg):
            raise RuntimeError('{} is not a qreg'.format(repr(qreg)))
        self.qreg = qreg

    def check_constraints(self):
        # gate_type.py:7-10
        if self.ctrllist is not None and self.qreg in self.ctrllist:
            raise RuntimeError('control and operand overlapped.')

    def copy(self):
        dup = Gate(self.gate_type)
        dup.adjoint = self.adjoint
        dup.qreg = self.qreg
        dup.ctrllist = None if self.ctrllist is None else list(self.ctrllist)
        return dup

    def __repr__(self):
        head = repr(self.gate_type) + ('.Adj' if self.adjoint else '')
        if self.ctrllist:
            head = 'ctrl({}).'.format(self.ctrllist) + head
        return '{}({})'.format(head, self.qreg)


class GatelistMacro(Operator):
    """Composite gate defined over a list of Pauli gates (Expi), model.py:90-123."""

    def __init__(self, gate_type):
        self.gate_type = gate_type
        self.adjoint = False
        self.gatelist = None
        self.ctrllist = None

    def set_adjoint(self, adjoint):
        self.adjoint = adjoint

    def set_ctrllist(self, ctrllist):
        self.ctrllist = _as_qreg_list(ctrllist)

    def set_gatelist(self, gatelist):
        self.gatelist = list(gatelist)

    def check_constraints(self):
        # gate_type.py:13-26
        targets = set()
        for gate in self.gatelist:
            if not isinstance(gate.gate_type, gate_type.PAULI):
                raise RuntimeError('exp gate only accepts ID, X, Y and Z gates')
            if gate.ctrllist is not None:
                raise RuntimeError('control qreg(s) should not be set for exp gate parameters.')
            targets.add(gate.qreg)
        if self.ctrllist is not None and targets & set(self.ctrllist):
            raise RuntimeError('control bit and target should not overlap.')

    def copy(self):
        dup = GatelistMacro(self.gate_type)
        dup.adjoint = self.adjoint
        dup.ctrllist = None if self.ctrllist is None else list(self.ctrllist)
        dup.gatelist = [gate.copy() for gate in self.gatelist]
        return dup


class MultiQubitGate(Operator):
    """Swap (model.py:126-146)."""

    def __init__(self, gate_type):
        self.gate_type = gate_type
        self.adjoint = False
        self.qreglist = None

    def set_adjoint(self, adjoint):
        self.adjoint = adjoint

    def set_qreglist(self, qreglist):
        self.qreglist = _as_qreg_list(qreglist)

    def check_constraints(self):
        pass

    def copy(self):
        dup = MultiQubitGate(self.gate_type)
        dup.adjoint = self.adjoint
        dup.qreglist = list(self.qreglist)
        return dup


class _QregObserver(Operator):
    def __init__(self, ref, qreg):
        if not isinstance(qreg, Qreg) or not isinstance(ref, Reference):
            raise RuntimeError('Wrong argument for {}, {}, {}.'.format(
                type(self).__name__, repr(qreg), repr(ref)))
        self.qreg, self.outref = qreg, ref

    def copy(self):
        return type(self)(self.outref, self.qreg)

    def __repr__(self):
        return '{}({}, {})'.format(type(self).__name__, self.outref, self.qreg)


class Measure(_QregObserver):
    """Z-basis measurement of one qreg (model.py:148-157)."""


class Prob(_QregObserver):
    """P(qreg == 0) into a reference (model.py:159-168)."""


class _PauliObserver(Operator):
    def __init__(self, ref, gatelist):
        if not isinstance(ref, Reference):
            raise RuntimeError('Wrong argument for {}, {}.'.format(type(self).__name__, repr(ref)))
        for gate in gatelist:
            if not isinstance(gate.gate_type, gate_type.PAULI):
                raise RuntimeError('Pmeasure only accepts ID, X, Y and Z gates')
            if gate.ctrllist is not None:
                raise RuntimeError('control qreg(s) should not be set for pauli operators.')
        self.gatelist, self.outref = list(gatelist), ref

    def copy(self):
        return type(self)(self.outref, [gate.copy() for gate in self.gatelist])


class PauliMeasure(_PauliObserver):
    """Measurement of a Pauli-string observable (model.py:170-191)."""


class PauliProb(_PauliObserver):
    """Probability of the +1 eigenspace of a Pauli string (model.py:193-199)."""


class Barrier(Operator):
    def __init__(self, qregset):
        self.qregset = set(_as_qreg_list(qregset))


class Reset(Operator):
    def __init__(self, qreg):
        if not isinstance(qreg, Qreg):
            raise RuntimeError('{} is not a qreg'.format(repr(qreg)))
        self.qreg = qreg


class IfClause(Operator):
    def __init__(self, refs, cond, clause):
        self.refs = refs
        self.cond = cond
        self.clause = clause


# --- directives inserted by the preprocessor (qgate/model/directive.py) ---------------
class NewQreg(Operator):
    def __init__(self, qreg):
        self.qreg = qreg


class ReleaseQreg(Operator):
    def __init__(self, qreg):
        self.qreg = qreg


class Join(Operator):
    def __init__(self, qregs):
        self.qreglist = list(qregs)
        assert 1 < len(self.qreglist)


class Separate(Operator):
    def __init__(self, qreg):
        self.qreg = qreg


# --- gate list -------------------------------------------------------------------------
class GateList:
    """Nested operator sequence; nested python lists become nested GateLists
    (qgate/model/gatelist.py:4-80)."""

    def __init__(self, ops=None):
        self.ops = []
        if ops is not None:
            self.set(ops)

    @staticmethod
    def _import(ops):
        if isinstance(ops, Operator):
            return [ops.copy()]
        if isinstance(ops, GateList):
            ops = ops.ops
        imported = []
        for op in ops:
            if isinstance(op, Operator):
                imported.append(op.copy())
            elif isinstance(op, (list, tuple, GateList)):
                inner = GateList()
                inner.ops = GateList._import(op)
                imported.append(inner)
            else:
                raise RuntimeError('Unknown argument, {}'.format(repr(op)))
        return imported

    def set(self, ops):
        self.ops = GateList._import(ops)

    def copy(self):
        return GateList(self)

    def append(self, op):
        self.ops += GateList._import(op)

    def __iadd__(self, other):
        self.ops += GateList._import(other)
        return self

    def __add__(self, other):
        merged = GateList(self)
        merged += other
        return merged

    def __iter__(self):
        return iter(self.ops)

    def __getitem__(self, key):
        return self.ops[key]

    def __len__(self):
        return len(self.ops)


def flatten(ops):
    """Depth-first iteration over nested GateLists / lists, yielding operators."""
    for op in (ops.ops if isinstance(ops, GateList) else ops):
        if isinstance(op, (GateList, list, tuple)):
            for inner in flatten(op):
                yield inner
        else:
            yield op
