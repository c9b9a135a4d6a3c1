"""`Qubits`: the user's view of all state vectors of a simulator — ordering, amplitude and
probability slicing, per-qreg probability, sampling pools.  Behavioural mirror of
qgate/simulator/qubits.py:6-214 and lanes.py:3-51 on top of the runtime getter protocol.
"""
import weakref

import numpy as np

from .. import model
from ..observation import ObservationList


# math operations selectable in get_states (qubits.py:6-10)
def null(v):
    return v


def abs2(v):
    return v.real ** 2 + v.imag ** 2


prob = abs2


class Lane:
    """A qreg's bit position: `local` inside its qstates, `external` in the user ordering."""
    __slots__ = ('qstates', 'local', 'external')

    def __init__(self, qstates, local, external=None):
        self.qstates = qstates
        self.local = local
        self.external = external


class Lanes(dict):
    def add_lane(self, qreg, qstates, local):
        self[qreg] = Lane(qstates, local)

    def get_by_qubit_states(self, qstates):
        return [lane for lane in self.values() if lane.qstates is qstates]


def create_lane_transformation(lanes, qreg_ordering):
    """[(qstates, [Lane(local, external) sorted by local]) ...]; external = -1 for qregs that
    are not in the ordering (lanes.py:30-51)."""
    position = {qreg: idx for idx, qreg in enumerate(qreg_ordering)}
    per_qstates = dict()
    for qreg, lane in lanes.items():
        per_qstates.setdefault(lane.qstates, []).append(
            Lane(qreg, lane.local, position.get(qreg, -1)))
    transformation = []
    for qstates, lanelist in per_qstates.items():
        lanelist.sort(key=lambda lane: lane.local)
        transformation.append((qstates, lanelist))
    return transformation


class EmptySamplingPool:
    """Pool over no allocated qreg: every sample is 0 (empty_sampling_pool.py:4-11)."""

    def __init__(self, qreg_ordering):
        self.qreg_ordering = qreg_ordering

    def sample(self, n_samples, randnum=None):
        return ObservationList(self.qreg_ordering, np.zeros([n_samples], np.int64), 0)


class StateGetter:
    def __init__(self, qubits, mathop, n_qregs):
        self._qubits = weakref.ref(qubits)
        self._mathop = mathop
        self._n_qregs = n_qregs

    @property
    def n_qregs(self):
        return self._n_qregs

    @property
    def mathop(self):
        return self._mathop

    def __getitem__(self, key):
        qubits = self._qubits()
        if qubits is None:
            return None
        return qubits.get_states(self._mathop, key)


class Qubits:
    def __init__(self, states_getter, dtype):
        self.states_getter = states_getter
        self.dtype = dtype
        self._given_ordering = None
        self._ordering = []
        self.lanes = Lanes()
        self.qstates_list = []

    def reset(self):
        self.lanes.clear()
        self.qstates_list = []

    def get_n_lanes(self):
        return len(self.lanes)

    def get_n_qregs(self):
        qregset = set(self.lanes.keys())
        if self._given_ordering is not None:
            qregset |= set(self._given_ordering)
        return len(qregset)

    @property
    def ordering(self):
        return self._ordering

    def set_ordering(self, qreglist):
        self._given_ordering = list(qreglist)
        self.update_external_layout()

    def update_external_layout(self):
        given = list(self._given_ordering) if self._given_ordering is not None else []
        rest = sorted(set(self.lanes.keys()) - set(given), key=lambda qreg: qreg.id)
        self._ordering = given + rest

    @property
    def states(self):
        return StateGetter(self, null, self.get_n_qregs())

    @property
    def prob(self):
        return StateGetter(self, abs2, self.get_n_qregs())

    def calc_probability(self, qreg):
        if not isinstance(qreg, model.Qreg):
            raise RuntimeError('qreg must be an instance of class Qreg.')
        lane = self.lanes[qreg]
        return lane.qstates.calc_probability(lane.local)

    def create_sampling_pool(self, qreg_ordering, sampling_pool_factory=None):
        qreg_ordering = list(qreg_ordering)
        if len(set(qreg_ordering)) != len(qreg_ordering):
            raise RuntimeError('qreg_ordering has duplicate qregs, {}.'.format(repr(qreg_ordering)))
        pool_ordering = [qreg for qreg in qreg_ordering if qreg in self.lanes]
        empty_lanes = [pos for pos, qreg in enumerate(qreg_ordering) if qreg not in self.lanes]
        if len(pool_ordering) == 0:
            return EmptySamplingPool(qreg_ordering)
        lane_trans = create_lane_transformation(self.lanes, pool_ordering)
        n_hidden_lanes = len(self.lanes) - len(pool_ordering)
        return self.states_getter.create_sampling_pool(qreg_ordering, len(pool_ordering),
                                                       n_hidden_lanes, lane_trans, empty_lanes,
                                                       sampling_pool_factory)

    def get_states(self, mathop=null, key=None):
        if mathop is null:
            dtype = np.complex64 if np.dtype(self.dtype) == np.float32 else np.complex128
        elif mathop is abs2:
            dtype = np.dtype(self.dtype).type
        else:
            raise RuntimeError('unsupported mathop, {}'.format(repr(mathop)))

        empty_lanes = []
        if self._given_ordering is not None:
            empty_lanes = [pos for pos, qreg in enumerate(self._given_ordering)
                           if qreg not in self.lanes]
        lane_trans = create_lane_transformation(self.lanes, self._ordering)
        n_states = 1 << self.get_n_qregs()

        if key is None:
            key = slice(0, n_states)
        if isinstance(key, slice):
            # python list slicing semantics (clipping, negative indices and steps),
            # like qubits.py:151-190
            span = range(*key.indices(n_states))
            if len(span) == 0:
                return np.ones([0], dtype)
            values = np.empty([len(span)], dtype)
            self.states_getter.get_states(values, 0, mathop, lane_trans, empty_lanes,
                                          len(span), span.start, span.step)
            return values

        idx = int(key)
        if idx < 0:
            if idx < -n_states:
                raise ValueError('list index out of range')
            idx += n_states
        if n_states <= idx:
            raise RuntimeError('list index out of range')
        values = np.empty([1], dtype)
        self.states_getter.get_states(values, 0, mathop, lane_trans, empty_lanes, 1, idx, 1)
        return values[0]
