"""`dump(obj)` pretty printer (qgate/__init__.py:6-28, qgate/simulator/utils.py)."""
import sys

from . import model
from .observation import ObservationHistgram, ObservationList


def dump(obj, file=None, number_format=None):
    from .simulator import qubits as _qubits
    from .simulator.simulator import ValueStore
    file = sys.stdout if file is None else file
    if isinstance(obj, (model.GateList, list)):
        for idx, op in enumerate(model.flatten(obj)):
            print('{}: {}'.format(idx, repr(op)), file=file)
    elif isinstance(obj, (_qubits.Qubits, _qubits.StateGetter)):
        getter = obj.states if isinstance(obj, _qubits.Qubits) else obj
        if number_format is None:
            number_format = '{1: .3f}' if getter.mathop is _qubits.null else '{1:g}'
        fmt = '|{{0:0{0:d}b}}> '.format(max(1, getter.n_qregs)) + number_format
        for idx, value in enumerate(getter[:]):
            print(fmt.format(idx, value), file=file)
    elif isinstance(obj, ValueStore):
        for key, value in obj.valuedict.items():
            print('{:d}:'.format(key), value, file=file)
    elif isinstance(obj, ObservationList):
        for observation in obj:
            print(repr(observation), file=file)
    elif isinstance(obj, ObservationHistgram):
        print(repr(obj), file=file)
    else:
        raise RuntimeError('unknow object, {}.'.format(type(obj)))
