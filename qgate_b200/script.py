"""User-facing circuit-building API, name-compatible with `qgate.script`
(qgate/script/script.py:18-310): new_qregs, new_references, H, X, ..., U3,
ctrl(...).X, measure, prob, barrier, reset, if_, Swap, Expi.

    from qgate_b200.script import *
    q = new_qregs(3)
    circuit = [H(q[0]), ctrl(q[0]).X(q[1]), ctrl(q[0], q[1]).U1(0.3).Adj(q[2])]
"""
import numbers

from . import model
from .model import gate_type as gtype

__all__ = ['new_qreg', 'new_qregs', 'release_qreg', 'new_reference', 'new_references',
           'new_gatelist', 'measure', 'prob', 'barrier', 'reset', 'if_',
           'I', 'H', 'S', 'T', 'X', 'Y', 'Z', 'Rx', 'Ry', 'Rz', 'U1', 'U2', 'U3',
           'controlled', 'ctrl', 'Swap', 'SH', 'Expii', 'Expiz', 'Expi']


def _flatten_args(args):
    if isinstance(args, (list, tuple, set)):
        flat = []
        for child in args:
            flat += _flatten_args(child)
        return flat
    return [args]


def _numbers(*values):
    for value in values:
        if not isinstance(value, numbers.Number):
            raise RuntimeError('{} is not a number'.format(str(value)))
    return values


def new_gatelist():
    return model.GateList()


def new_qreg():
    return model.Qreg()


def new_qregs(count):
    return [model.Qreg() for _ in range(count)]


def release_qreg(qreg):
    return model.ReleaseQreg(qreg)


def new_reference():
    return model.Reference()


def new_references(count):
    return [model.Reference() for _ in range(count)]


def measure(outref, args):
    if isinstance(args, model.Qreg):
        return model.Measure(outref, args)
    return model.PauliMeasure(outref, _flatten_args(args))


def prob(outref, args):
    if isinstance(args, model.Qreg):
        return model.Prob(outref, args)
    return model.PauliProb(outref, _flatten_args(args))


def barrier(*qregs):
    return model.Barrier(_flatten_args(qregs))


def reset(qreg):
    return model.Reset(qreg)


def if_(refs, cond, clause):
    return model.IfClause(_flatten_args(refs), cond, model.GateList(clause))


class _GateBuilder:
    """`H`, `U3(a,b,c)`, `ctrl(q).X` ... : call with the target qreg to get the Gate;
    `.Adj` selects the adjoint."""

    def __init__(self, gate_type, ctrllist=None, adjoint=False):
        self._gate_type = gate_type
        self._ctrllist = ctrllist
        self._adjoint = adjoint

    @property
    def Adj(self):
        return _GateBuilder(self._gate_type, self._ctrllist, True)

    def __call__(self, qreg):
        gate = model.Gate(self._gate_type)
        gate.set_adjoint(self._adjoint)
        if self._ctrllist is not None:
            gate.set_ctrllist(self._ctrllist)
        gate.set_qreg(qreg)
        gate.check_constraints()
        return gate


class _MacroBuilder:
    def __init__(self, gate_type, ctrllist=None, adjoint=False):
        self._gate_type = gate_type
        self._ctrllist = ctrllist
        self._adjoint = adjoint

    @property
    def Adj(self):
        return _MacroBuilder(self._gate_type, self._ctrllist, True)

    def __call__(self, *gatelist):
        macro = model.GatelistMacro(self._gate_type)
        macro.set_adjoint(self._adjoint)
        if self._ctrllist is not None:
            macro.set_ctrllist(self._ctrllist)
        macro.set_gatelist(_flatten_args(gatelist))
        macro.check_constraints()
        return macro


_CONST_GATES = {'I': gtype.ID, 'H': gtype.H, 'S': gtype.S, 'T': gtype.T, 'X': gtype.X,
                'Y': gtype.Y, 'Z': gtype.Z, 'SH': gtype.SH}
_PARAM_GATES = {'Rx': gtype.RX, 'Ry': gtype.RY, 'Rz': gtype.RZ, 'U1': gtype.U1, 'U2': gtype.U2,
                'U3': gtype.U, 'Expii': gtype.ExpiI, 'Expiz': gtype.ExpiZ}

I = _GateBuilder(gtype.ID())
H = _GateBuilder(gtype.H())
S = _GateBuilder(gtype.S())
T = _GateBuilder(gtype.T())
X = _GateBuilder(gtype.X())
Y = _GateBuilder(gtype.Y())
Z = _GateBuilder(gtype.Z())
SH = _GateBuilder(gtype.SH())


def Rx(theta):
    return _GateBuilder(gtype.RX(*_numbers(theta)))


def Ry(theta):
    return _GateBuilder(gtype.RY(*_numbers(theta)))


def Rz(theta):
    return _GateBuilder(gtype.RZ(*_numbers(theta)))


def U1(_lambda):
    return _GateBuilder(gtype.U1(*_numbers(_lambda)))


def U2(phi, _lambda):
    return _GateBuilder(gtype.U2(*_numbers(phi, _lambda)))


def U3(theta, phi, _lambda):
    return _GateBuilder(gtype.U(*_numbers(theta, phi, _lambda)))


def Expii(theta):
    return _GateBuilder(gtype.ExpiI(*_numbers(theta)))


def Expiz(theta):
    return _GateBuilder(gtype.ExpiZ(*_numbers(theta)))


def Expi(theta):
    return _MacroBuilder(gtype.Expi(theta))


class _Controlled:
    """ctrl(q0, q1, ...).<gate> — multi-controlled single-qubit gates."""

    def __init__(self, control):
        self._control = _flatten_args(control)
        if len(self._control) == 0:
            raise RuntimeError('control qreg list must not be empty.')

    def __getattr__(self, name):
        if name in _CONST_GATES:
            return _GateBuilder(_CONST_GATES[name](), self._control)
        if name in _PARAM_GATES:
            cls = _PARAM_GATES[name]

            def factory(*params):
                return _GateBuilder(cls(*_numbers(*params)), self._control)
            return factory
        if name == 'Expi':
            return lambda theta: _MacroBuilder(gtype.Expi(theta), self._control)
        raise AttributeError(name)


def controlled(*control):
    return _Controlled(control)


ctrl = controlled


def Swap(qreg0, qreg1):
    gate = model.MultiQubitGate(gtype.SWAP())
    gate.set_qreglist([qreg0, qreg1])
    return gate
