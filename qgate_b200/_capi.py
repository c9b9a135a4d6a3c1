"""ctypes binding of include/qgate_b200.h.

One `CApi` instance wraps one shared library implementing the header.  The
product uses exactly one: qgate_b200/lib/libqgate_b200.so (CUDA, sm_100a).
Tests additionally bind oracle/_ref/libqgate_ref_cpu.so (the unmodified
reference CPU runtime behind the same ABI) through the same class.

This replaces the CPython argument unpacking of the reference's glue.cpp
(qgate/simulator/src/glue.cpp:232-602): no native object ever sees a Python
type; errors come back as status codes and are raised here.
"""
import ctypes as C
import os

import numpy as np

OK, ERR_INVALID, ERR_RUNTIME, ERR_OOM, ERR_CUDA = 0, 1, 2, 3, 4
PREC_FP64, PREC_FP32 = 1, 2
MATHOP_NULL, MATHOP_PROB = 0, 1

GATE_IDS = {
    'U': 0, 'U2': 1, 'U1': 2, 'ID': 3, 'X': 4, 'Y': 5, 'Z': 6, 'H': 7, 'S': 8, 'T': 9,
    'RX': 10, 'RY': 11, 'RZ': 12, 'ExpiI': 13, 'ExpiZ': 14, 'SH': 15,
}

_h = C.c_uint64
_i = C.c_int
_i64 = C.c_int64
_d = C.c_double
_p = C.c_void_p
_ip = C.POINTER(C.c_int)
_hp = C.POINTER(C.c_uint64)
_dp = C.POINTER(C.c_double)
_i64p = C.POINTER(C.c_int64)


class Stats(C.Structure):
    _fields_ = [(name, C.c_int64) for name in (
        'kernel_launches', 'gates_submitted', 'gates_executed', 'tile_passes',
        'gate_amp_updates', 'pass_bytes', 'h2d_bytes', 'd2h_bytes', 'tma_passes', 'shear_ops',
        'direct_ops', 'native_swaps', 'native_pauli_exps', 'fan_ops')]

    def as_dict(self):
        return {name: int(getattr(self, name)) for name, _ in self._fields_}


# name -> argtypes; every function returns int status unless listed in _NON_STATUS.
SIGNATURES = {
    'qgb_devices_initialize': [_ip, _i, _i, _i64],
    'qgb_devices_clear': [],
    'qgb_device_count': [_ip],
    'qgb_set_stream': [C.c_uint64],
    'qgb_gate_matrix': [_i, _dp, _i, _i, _dp],
    'qgb_qstates_new': [_i, _hp],
    'qgb_qstates_delete': [_h],
    'qgb_qstates_deallocate': [_h],
    'qgb_qstates_get_n_lanes': [_h, _ip],
    'qgb_qproc_new': [_i, _hp],
    'qgb_qproc_delete': [_h],
    'qgb_qproc_synchronize': [_h],
    'qgb_qproc_reset': [_h],
    'qgb_qproc_initialize_qstates': [_h, _h, _i],
    'qgb_qproc_reset_qstates': [_h, _h],
    'qgb_qproc_calc_probability': [_h, _h, _i, _dp],
    'qgb_qproc_join': [_h, _h, _hp, _i, _i],
    'qgb_qproc_decohere': [_h, _i, _d, _h, _i],
    'qgb_qproc_decohere_and_separate': [_h, _i, _d, _h, _h, _h, _i],
    'qgb_qproc_apply_reset': [_h, _h, _i],
    'qgb_qproc_apply_gate': [_h, _dp, _h, _i],
    'qgb_qproc_apply_controlled_gate': [_h, _dp, _h, _ip, _i, _i],
    'qgb_qproc_apply_gate_typed': [_h, _i, _dp, _i, _i, _h, _ip, _i, _i],
    'qgb_getter_new': [_i, _hp],
    'qgb_getter_delete': [_h],
    'qgb_getter_get_states': [_h, _p, _i64, _i, _ip, _ip, _i64, _hp, _i, _i, _i64, _i64, _i64],
    'qgb_getter_prepare_prob_array': [_h, _p, _ip, _ip, _hp, _i, _i, _i],
    'qgb_getter_create_sampling_pool': [_h, _ip, _ip, _hp, _i, _i, _i, _ip, _i, _hp],
    'qgb_pool_sample': [_h, _i64p, _i, _dp],
    'qgb_pool_delete': [_h],
    'qgb_pool_sample_sequential': [_h, _i64p, _i, _dp],
    'qgb_stats_get': [C.POINTER(Stats)],
    'qgb_stats_reset': [],
    'qgb_qproc_flush': [_h, _h],
    'qgb_qproc_apply_swap': [_h, _h, _i, _i],
    'qgb_qproc_apply_pauli_expi': [_h, _h, _d, _ip, _ip, _i, _ip, _i],
    'qgb_qproc_apply_gates_batch': [_h, _h, _p, _i64],
    'qgb_set_option': [C.c_char_p, _i64],
    # sharding support
    'qgb_qstates_data_ptr': [_h, _hp, _i64p],
    'qgb_qstates_alt_buffer': [_h, _hp],
    'qgb_qstates_flip': [_h],
    'qgb_qproc_calc_norm': [_h, _h, _dp],
    'qgb_qproc_join_shard': [_h, _h, _hp, _i, _i, _i, _i64],
    'qgb_getter_create_sampling_pool_partial': [_h, _ip, _ip, _hp, _i, _i, _i, _hp, _dp],
    'qgb_pool_finalize': [_h, _d, _d],
    'qgb_pool_from_prob_array': [_i, _dp, _i, _ip, _i, _hp],
    'qgb_qstates_ipc_export': [_h, _p, _i64p],
    'qgb_ipc_open': [_p, _hp],
    'qgb_ipc_close': [C.c_uint64],
    'qgb_qstates_exchange_p2p': [_h, _hp, _i, _ip, _i],
    'qgb_qstates_ipc_export_alt': [_h, _p, _i64p],
    'qgb_qstates_exchange_push': [_h, _hp, _i, _ip, _i],
    'qgb_pool_trim_exported': [],
}
_NON_STATUS = {
    'qgb_last_error': ([], C.c_char_p),
    'qgb_backend_name': ([], C.c_char_p),
    'qgb_abi_version': ([], C.c_int),
}


class GateOp(C.Structure):
    """qgb_gate_op of include/qgate_b200.h (one record of qgb_qproc_apply_gates_batch)."""
    _fields_ = [('gate_id', C.c_int32), ('adjoint', C.c_int32), ('target', C.c_int32), ('reserved', C.c_int32),
                ('ctrl_mask', C.c_uint64), ('args', C.c_double * 3), ('mat8', C.c_double * 8)]


GATE_OP_DTYPE = np.dtype([('gate_id', np.int32), ('adjoint', np.int32), ('target', np.int32), ('reserved', np.int32),
                          ('ctrl_mask', np.uint64), ('args', np.float64, (3,)), ('mat8', np.float64, (8,))])
assert GATE_OP_DTYPE.itemsize == C.sizeof(GateOp)
GATE_MATRIX = -1
PAULI_CODES = {'ID': 0, 'X': 1, 'Y': 2, 'Z': 3}


def exported_symbols():
    """Every symbol include/qgate_b200.h declares (used by the CPU-side ABI test)."""
    return sorted(list(SIGNATURES) + list(_NON_STATUS))


def prec_of(dtype):
    """np.float32 / np.float64 (classes, dtypes or names) -> QGB_PREC_*.

    The reference matches the NumPy scalar *type objects* by identity
    (src/pyglue.h:129-138); dtype instances are accepted here as a superset."""
    dt = np.dtype(dtype)
    if dt == np.float64:
        return PREC_FP64
    if dt == np.float32:
        return PREC_FP32
    raise RuntimeError('unsupported dtype, {}.'.format(repr(dtype)))


class CApi:
    def __init__(self, libpath):
        if not os.path.exists(libpath):
            raise ImportError('native library not found: {}'.format(libpath))
        self.libpath = libpath
        self.lib = C.CDLL(libpath, mode=C.RTLD_GLOBAL)
        for name, argtypes in SIGNATURES.items():
            fn = getattr(self.lib, name)
            fn.argtypes = argtypes
            fn.restype = C.c_int
        for name, (argtypes, restype) in _NON_STATUS.items():
            fn = getattr(self.lib, name)
            fn.argtypes = argtypes
            fn.restype = restype
        self.backend_name = self.lib.qgb_backend_name().decode()

    # -- error plumbing ------------------------------------------------------
    def check(self, status):
        if status == OK:
            return
        msg = self.lib.qgb_last_error().decode(errors='replace')
        if status == ERR_INVALID:
            raise ValueError(msg)
        if status == ERR_OOM:
            raise RuntimeError('Out of device memory. ' + msg)
        raise RuntimeError(msg)

    def call(self, name, *args):
        self.check(getattr(self.lib, name)(*args))

    # -- small helpers ---------------------------------------------------------
    @staticmethod
    def int_array(values):
        arr = (C.c_int * max(1, len(values)))()
        for idx, v in enumerate(values):
            arr[idx] = int(v)
        return arr

    @staticmethod
    def handle_array(handles):
        arr = (C.c_uint64 * max(1, len(handles)))()
        for idx, v in enumerate(handles):
            arr[idx] = int(v)
        return arr

    def new_handle(self, fname, prec):
        out = C.c_uint64(0)
        self.call(fname, prec, C.byref(out))
        return out.value

    def gate_matrix(self, gate_id, args, adjoint):
        cargs = (C.c_double * max(1, len(args)))(*[float(a) for a in args])
        mat = (C.c_double * 8)()
        self.call('qgb_gate_matrix', gate_id, cargs, len(args), 1 if adjoint else 0, mat)
        m = np.array(mat[:], np.float64)
        return (m[0::2] + 1j * m[1::2]).reshape(2, 2)

    def stats(self):
        st = Stats()
        self.call('qgb_stats_get', C.byref(st))
        return st.as_dict()

    def stats_reset(self):
        self.call('qgb_stats_reset')

    def set_option(self, name, value):
        self.call('qgb_set_option', name.encode(), int(value))

    def set_stream(self, cuda_stream):
        self.call('qgb_set_stream', int(cuda_stream))

    def device_count(self):
        n = C.c_int(0)
        self.call('qgb_device_count', C.byref(n))
        return n.value
