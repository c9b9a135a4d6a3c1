"""Host-side object protocol of a native runtime.

`Simulator`, `Qubits`, `QubitsHandler` and the executors talk to a runtime
through these four objects and the module-level functions of `RuntimeModule`.
Names, arguments and error behaviour mirror the reference wrappers:

  NativeQubitStates        qgate/simulator/native_qubit_states.py:4-33
  NativeQubitProcessor     qgate/simulator/native_qubit_processor.py:6-59
  NativeQubitsStatesGetter qgate/simulator/native_qubits_states_getter.py:6-102
  NativeSamplingPool       qgate/simulator/native_sampling_pool.py:5-24
  RuntimeModule            qgate/simulator/cudaruntime.py:24-91

The classes accept gate-type / lane / qreg objects from either this package's
front end or the reference's own (duck-typed: `type(gate_type).__name__`,
`.args`, `.local`, `.external`), so the same module can be installed as
`qgate.simulator.cudaruntime` (see qgate_b200/install.py).
"""
import ctypes as C
import weakref

import numpy as np

from . import _capi
from .observation import ObservationList


def _mathop_id(mathop):
    # qgate/simulator/native_qubits_states_getter.py:21-26 compares function objects;
    # match by name so both this package's and the reference's qubits.null/abs2 work.
    name = getattr(mathop, '__name__', None)
    if name == 'null':
        return _capi.MATHOP_NULL
    if name in ('abs2', 'prob'):
        return _capi.MATHOP_PROB
    raise RuntimeError('unknown math operation, {}'.format(str(mathop)))


class NativeQubitStates:
    def __init__(self, api, ptr, processor):
        self.api = api
        self.ptr = ptr
        self.processor = processor

    def __del__(self):
        try:
            self.delete()
        except Exception:
            pass

    def delete(self):
        if hasattr(self, 'ptr'):
            self.api.call('qgb_qstates_delete', self.ptr)
            del self.ptr
        if hasattr(self, 'processor'):
            self.processor.delete()
            del self.processor

    def get_n_lanes(self):
        n = C.c_int(0)
        self.api.call('qgb_qstates_get_n_lanes', self.ptr, C.byref(n))
        return n.value

    # measured value per lane lives on the host only (native_qubit_states.py:23-30)
    def reset_lane_states(self):
        self.lane_states = [-1] * self.get_n_lanes()

    def get_lane_state(self, lane):
        return self.lane_states[lane]

    def set_lane_state(self, lane, value):
        self.lane_states[lane] = value

    def calc_probability(self, lane):
        return self.processor.calc_probability(self, lane)


class NativeQubitProcessor:
    def __init__(self, api, dtype, ptr):
        self.api = api
        self.dtype = dtype
        self.ptr = ptr
        self._lib = api.lib
        self._check = api.check

    def __del__(self):
        try:
            self.delete()
        except Exception:
            pass

    def delete(self):
        if hasattr(self, 'ptr'):
            self.api.call('qgb_qproc_delete', self.ptr)
            del self.ptr

    def synchronize(self):
        self.api.call('qgb_qproc_synchronize', self.ptr)

    def reset(self):
        self.api.call('qgb_qproc_reset', self.ptr)

    def initialize_qubit_states(self, qstates, n_lanes):
        self.api.call('qgb_qproc_initialize_qstates', self.ptr, qstates.ptr, n_lanes)
        qstates.reset_lane_states()

    def reset_qubit_states(self, qstates):
        self.api.call('qgb_qproc_reset_qstates', self.ptr, qstates.ptr)
        qstates.reset_lane_states()

    def calc_probability(self, qstates, local_lane):
        prob = C.c_double(0.)
        self.api.call('qgb_qproc_calc_probability', self.ptr, qstates.ptr, local_lane,
                      C.byref(prob))
        return prob.value

    def join(self, qstates, qstates_list, n_new_qregs):
        ptrs = self.api.handle_array([qs.ptr for qs in qstates_list])
        self.api.call('qgb_qproc_join', self.ptr, qstates.ptr, ptrs, len(qstates_list),
                      n_new_qregs)

    def decohere(self, value, prob, qstates, local_lane):
        self.api.call('qgb_qproc_decohere', self.ptr, int(value), float(prob), qstates.ptr,
                      local_lane)

    def decohere_and_separate(self, value, prob, qstates0, qstates1, qstates, local_lane):
        self.api.call('qgb_qproc_decohere_and_separate', self.ptr, int(value), float(prob),
                      qstates0.ptr, qstates1.ptr, qstates.ptr, local_lane)

    def apply_reset(self, qstates, local_lane):
        self.api.call('qgb_qproc_apply_reset', self.ptr, qstates.ptr, local_lane)

    @staticmethod
    def _gate_key(gate_type):
        # (gate id, ctypes args, n_args) cached on the gate-type object: the host makes
        # exactly one C call per gate and builds no NumPy/ctypes temporaries for it.
        key = getattr(gate_type, '_qgb_key', None)
        if key is None:
            name = type(gate_type).__name__
            gate_id = _capi.GATE_IDS.get(name)
            if gate_id is None:
                raise RuntimeError('Unknown gate type.')
            args = [float(a) for a in gate_type.args]
            cargs = (C.c_double * max(1, len(args)))(*args)
            key = (gate_id, cargs, len(args))
            try:
                gate_type._qgb_key = key
            except AttributeError:
                pass
        return key

    def apply_gate(self, gate_type, _adjoint, qstates, local_lane):
        gate_id, cargs, n_args = self._gate_key(gate_type)
        rc = self._lib.qgb_qproc_apply_gate_typed(self.ptr, gate_id, cargs, n_args,
                                                  1 if _adjoint else 0, qstates.ptr,
                                                  None, 0, local_lane)
        if rc:
            self._check(rc)

    def apply_controlled_gate(self, gate_type, _adjoint, qstates, local_control_lanes,
                              local_target_lane):
        gate_id, cargs, n_args = self._gate_key(gate_type)
        n_ctrl = len(local_control_lanes)
        ctrl = (C.c_int * n_ctrl)(*local_control_lanes)
        rc = self._lib.qgb_qproc_apply_gate_typed(self.ptr, gate_id, cargs, n_args,
                                                  1 if _adjoint else 0, qstates.ptr,
                                                  ctrl, n_ctrl, local_target_lane)
        if rc:
            self._check(rc)

    def flush(self, qstates):
        self.api.call('qgb_qproc_flush', self.ptr, qstates.ptr)

    # -- "next" rows of the hot path: what the reference's front end expands
    #    or loops over in Python, as single native calls ------------------------------------------
    def apply_swap(self, qstates, lane_a, lane_b):
        """Swap (model/expand.py:14-20 expands it into three CX gates): a lane relabelling."""
        self.api.call('qgb_qproc_apply_swap', self.ptr, qstates.ptr, int(lane_a), int(lane_b))
        states = qstates.lane_states
        states[lane_a], states[lane_b] = states[lane_b], states[lane_a]

    def apply_pauli_expi(self, theta, qstates, lanes, paulis, ctrl_lanes=()):
        """exp(i theta P) for the Pauli string P = paulis[i] (0 I, 1 X, 2 Y, 3 Z) on lanes[i]
        (model/expand.py:35-51 expands it into basis changes + CX chains + ExpiZ)."""
        n, n_ctrl = len(lanes), len(ctrl_lanes)
        self.api.call('qgb_qproc_apply_pauli_expi', self.ptr, qstates.ptr, float(theta),
                      self.api.int_array(lanes), self.api.int_array(paulis), n,
                      self.api.int_array(ctrl_lanes), n_ctrl)

    @staticmethod
    def pack_gates(gates):
        """[(gate_type, adjoint, ctrl_lanes, target_lane)] -> the record array of
        qgb_qproc_apply_gates_batch (one native call for the whole list instead of one per gate,
        rop_executor.py:28-53)."""
        rec = np.zeros(len(gates), _capi.GATE_OP_DTYPE)
        for idx, (gate_type, adjoint, ctrl_lanes, target) in enumerate(gates):
            gate_id, cargs, n_args = NativeQubitProcessor._gate_key(gate_type)
            row = rec[idx]
            row['gate_id'] = gate_id
            row['adjoint'] = 1 if adjoint else 0
            row['target'] = target
            mask = 0
            for lane in ctrl_lanes:
                mask |= 1 << lane
            row['ctrl_mask'] = mask
            for k in range(n_args):
                row['args'][k] = cargs[k]
        return rec

    def apply_gates_batch(self, qstates, records):
        records = np.ascontiguousarray(records, _capi.GATE_OP_DTYPE)
        self.api.call('qgb_qproc_apply_gates_batch', self.ptr, qstates.ptr,
                      records.ctypes.data_as(C.c_void_p), int(records.size))


class NativeSamplingPool:
    def __init__(self, api, ptr, qreg_ordering, mask):
        self.api = api
        self.ptr = ptr
        self.qreg_ordering = qreg_ordering
        self.mask = mask

    def __del__(self):
        try:
            self.delete()
        except Exception:
            pass

    def delete(self):
        if hasattr(self, 'ptr'):
            self.api.call('qgb_pool_delete', self.ptr)
            del self.ptr

    def sample(self, n_samples, randnum=None):
        obs = np.empty([n_samples], np.int64)
        if randnum is None:
            randnum = np.random.random_sample([n_samples])
        randnum = np.ascontiguousarray(randnum, np.float64)
        if randnum.size < n_samples:
            raise ValueError('array size too small.')
        self.api.call('qgb_pool_sample', self.ptr,
                      obs.ctypes.data_as(C.POINTER(C.c_int64)), int(n_samples),
                      randnum.ctypes.data_as(C.POINTER(C.c_double)))
        return ObservationList(self.qreg_ordering, obs, self.mask)

    def sample_sequential(self, n_shots, randnum):
        """n_shots shots of sequential measurements of the pool's qregs, the LAST qreg of the
        ordering measured first; randnum[s, k] decides the k-th measurement of shot s exactly as
        the reference's Measure op would (0 if r < P(0 | earlier outcomes), model_executor.py:
        117-122).  Returns the pool indices (int64[n_shots]); SURVEY section 8f-2."""
        randnum = np.ascontiguousarray(randnum, np.float64)
        obs = np.empty([n_shots], np.int64)
        self.api.call('qgb_pool_sample_sequential', self.ptr, obs.ctypes.data_as(C.POINTER(C.c_int64)),
                      int(n_shots), randnum.ctypes.data_as(C.POINTER(C.c_double)))
        return obs


class NativeQubitsStatesGetter:
    def __init__(self, api, dtype, ptr):
        self.api = api
        self.dtype = dtype
        self.ptr = ptr

    def __del__(self):
        try:
            self.delete()
        except Exception:
            pass

    def delete(self):
        if hasattr(self, 'ptr'):
            self.api.call('qgb_getter_delete', self.ptr)
            del self.ptr

    # lane_trans = [(qstates, [Lane(.local, .external) sorted by local]) ...]  (lanes.py:30-51)
    def _pack_transform(self, lane_trans):
        ptrs, tables, counts = [], [], []
        for qstates, lanepos_list in lane_trans:
            ptrs.append(qstates.ptr)
            table = [None] * len(lanepos_list)
            for lanepos in lanepos_list:
                table[lanepos.local] = lanepos.external
            tables += table
            counts.append(len(table))
        return (self.api.handle_array(ptrs), self.api.int_array(tables),
                self.api.int_array(counts), len(ptrs))

    @staticmethod
    def create_lane_mask(lanepos_list):
        mask = 0
        for pos in lanepos_list:
            mask |= 1 << pos
        return mask

    def get_states(self, values, offset, mathop, lane_trans, empty_lanes, n_states, start, step):
        mathop = _mathop_id(mathop)
        n_lanes = sum([len(lanepos_list) for _, lanepos_list in lane_trans])
        n_qregs = n_lanes + len(empty_lanes)
        empty_lane_mask = self.create_lane_mask(empty_lanes)
        if not (isinstance(values, np.ndarray) and values.flags['C_CONTIGUOUS']):
            raise RuntimeError('values must be a contiguous ndarray.')
        if values.size - offset < n_states:
            raise ValueError('array size too small.')
        ptrs, tables, counts, n_qs = self._pack_transform(lane_trans)
        self.api.call('qgb_getter_get_states', self.ptr, values.ctypes.data_as(C.c_void_p),
                      int(offset), mathop, tables, counts, int(empty_lane_mask), ptrs, n_qs,
                      n_qregs, int(n_states), int(start), int(step))

    @staticmethod
    def place_hidden_lanes_in_lsb(lane_trans, n_lanes, n_hidden_lanes):
        # convention of the reference CPU runtime, native_qubits_states_getter.py:73-82:
        # contiguous groups of 2^n_hidden external indices are summed out.
        hidden_idx = 0
        for _, lanelist in lane_trans:
            for lane in lanelist:
                if lane.external == -1:
                    lane.external = hidden_idx
                    hidden_idx += 1
                else:
                    lane.external += n_hidden_lanes

    def create_sampling_pool(self, qreg_ordering, n_lanes, n_hidden_lanes, lane_trans,
                             empty_lanes, sampling_pool_factory=None):
        self.place_hidden_lanes_in_lsb(lane_trans, n_lanes, n_hidden_lanes)
        ptrs, tables, counts, n_qs = self._pack_transform(lane_trans)
        if sampling_pool_factory is not None:
            # test seam of the reference (native_qubits_states_getter.py:50-54):
            # the factory just receives the marginal probability vector.
            prob = np.empty([1 << n_lanes], self.dtype)
            self.api.call('qgb_getter_prepare_prob_array', self.ptr,
                          prob.ctypes.data_as(C.c_void_p), tables, counts, ptrs, n_qs,
                          n_lanes, n_hidden_lanes)
            return sampling_pool_factory(prob, empty_lanes, qreg_ordering)
        pool = C.c_uint64(0)
        self.api.call('qgb_getter_create_sampling_pool', self.ptr, tables, counts, ptrs, n_qs,
                      n_lanes, n_hidden_lanes, self.api.int_array(empty_lanes),
                      len(empty_lanes), C.byref(pool))
        return NativeSamplingPool(self.api, pool.value, qreg_ordering,
                                  self.create_lane_mask(empty_lanes))


class RuntimeModule:
    """Module-level protocol of a runtime (cudaruntime.py:24-91), bound to one CApi."""

    native_multi_qubit_ops = True    # apply_swap / apply_pauli_expi exist on the processors

    @property
    def native_sample(self):           # pools answer Simulator.sample for terminal measurements
        return self.api.backend_name.startswith('cuda')

    def __init__(self, api_factory):
        self._api_factory = api_factory
        self._api = None
        self.initialized = False
        self.native_instances = weakref.WeakValueDictionary()
        self.reset_preference()

    @property
    def api(self):
        if self._api is None:
            self._api = self._api_factory()
        return self._api

    def set_preference(self, device_ids=[], max_po2idx_per_chunk=-1, memory_store_size=-1):
        # semantics of the second (live) definition, cudaruntime.py:33-41
        if self.initialized:
            raise RuntimeError('already initialized.')
        if len(device_ids) != 0:
            self.device_ids = list(device_ids)
        if max_po2idx_per_chunk != -1:
            self.max_po2idx_per_chunk = max_po2idx_per_chunk
        if memory_store_size != -1:
            self.memory_store_size = memory_store_size

    def reset_preference(self, device_ids=[], max_po2idx_per_chunk=-1, memory_store_size=-1):
        self.device_ids = []
        self.max_po2idx_per_chunk = -1
        self.memory_store_size = -1

    def module_init(self):
        ids = self.api.int_array(self.device_ids)
        self.api.call('qgb_devices_initialize', ids, len(self.device_ids),
                      int(self.max_po2idx_per_chunk), int(self.memory_store_size))
        self.initialized = True

    def module_finalize(self):
        for obj in list(self.native_instances.values()):
            obj.delete()
        if self.initialized:
            self.api.call('qgb_devices_clear')
            # (the library has just closed every IPC mapping: forget the ones dist.py kept for re-use)
            import sys
            dist_module = sys.modules.get(__package__ + '.dist')
            if dist_module is not None:
                kept = dist_module._RETAINED.get(id(self.api.lib))
                if kept is not None:
                    kept['bases'].clear()
                    kept['key'] = None
        self.initialized = False

    def create_qubit_states(self, dtype):
        if not self.initialized:
            self.module_init()
        api = self.api
        prec = _capi.prec_of(dtype)
        qproc = NativeQubitProcessor(api, dtype, api.new_handle('qgb_qproc_new', prec))
        self.native_instances[id(qproc)] = qproc
        qstates = NativeQubitStates(api, api.new_handle('qgb_qstates_new', prec), qproc)
        self.native_instances[id(qstates)] = qstates
        return qstates

    def create_qubits_states_getter(self, dtype):
        api = self.api
        getter = NativeQubitsStatesGetter(api, dtype,
                                          api.new_handle('qgb_getter_new', _capi.prec_of(dtype)))
        self.native_instances[id(getter)] = getter
        return getter
