"""State vectors sharded over ranks: one process per GPU, `torch.distributed` for the plumbing.

    torchrun --nproc-per-node 8 script.py          # every rank runs the same script (SPMD)
        dist.init_process_group('nccl'); torch.cuda.set_device(local_rank)
        sim = qgate_b200.simulator.cuda(dtype=np.float64)      # picks this module up

Layout (the reference's own layout, MultiChunkPtr.h:22-26): a qstates of n lanes
with n >= shard_min_lanes keeps 2^(n-g) amplitudes per rank, g = log2(world size); the g
high-order PHYSICAL lanes are the rank number.  Every sharded qstates carries a map
logical lane -> physical lane, so lanes can change places without touching the host code above.

What needs no communication: gates on local lanes; controls on global lanes (the rank either
takes part in the gate or skips it); diagonal gates on global lanes (the rank picks its diagonal
entry and applies a phase).  A non-diagonal gate on a global lane first trades that lane for a
local one — an exchange of half the shard with the partner rank; several global lanes trade at
once as an all-to-all.  Which local lanes to give up is chosen with the whole deferred gate
queue in view (furthest next non-diagonal use first).  Two exchange engines:

  'p2p'         one CUDA kernel swapping amplitudes in place through peer pointers (CUDA IPC
                mappings of the other ranks' shards) over NVLink — qgate_b200/csrc/dist.cu;
  'collective'  torch.distributed batched isend/irecv into a spare buffer (NCCL on GPUs, gloo
                in the CPU tests); needs twice the memory and one more local copy.

Probabilities are local reductions + all_reduce; readout and sampling pools are served by the
owning ranks and combined.  The reference instead lets every kernel read and write remote
chunks through peer pointers and synchronises all devices after every gate
(CUDAQubitProcessor.cpp:47-50,270-285; ProcessorRelocator.cpp:51-73).

Small qstates (dynamic qubit grouping creates many) are replicated: every rank holds the same
copy and does the same work on it, so no communication is needed for them either.
"""
import ctypes as C
import math
import os
import weakref

import numpy as np
import torch
import torch.distributed as dist

from . import _capi
from .native import (NativeQubitProcessor, NativeQubitStates, NativeQubitsStatesGetter,
                     NativeSamplingPool, _mathop_id)
from .observation import ObservationList

_DIAG_ZERO = (2, 3, 4, 5)


class _Lane:
    __slots__ = ('local', 'external')

    def __init__(self, local, external):
        self.local = local
        self.external = external


class _CudaView:
    """Raw device memory as a torch tensor (through __cuda_array_interface__)."""

    def __init__(self, ptr, n_items, typestr):
        self.__cuda_array_interface__ = {'shape': (n_items,), 'typestr': typestr,
                                         'data': (int(ptr), False), 'version': 2}


def _tensor_view(ptr, nbytes, on_cuda, device):
    n = nbytes // 8
    if on_cuda:
        return torch.as_tensor(_CudaView(ptr, n, '<f8'), device=device)
    arr = np.ctypeslib.as_array((C.c_double * n).from_address(int(ptr)))
    return torch.from_numpy(arr)


class Gate:
    """One deferred gate in logical-lane coordinates."""
    __slots__ = ('mat', 'ctrls', 'target', 'diag', 'xmask', 'zmask')

    def __init__(self, mat, ctrls, target):
        self.mat = mat                       # ctypes c_double[8]
        self.ctrls = ctrls
        self.target = target
        self.diag = all(mat[i] == 0. for i in _DIAG_ZERO)
        # lanes the gate acts on non-diagonally / diagonally (a control acts as |1><1|): two gates
        # commute when, on every lane they share, both act diagonally
        self.xmask = 0 if self.diag else (1 << target)
        z = (1 << target) if self.diag else 0
        for c in ctrls:
            z |= 1 << c
        self.zmask = z


# IPC mappings of the peers' shards outlive the qstates that opened them (per library, process-wide:
# bench.py and run_configs.py go through several runtime modules): unmapping a 16 GiB buffer that has
# been written through costs ~0.1 s (profiles/r2zd), and the engine's pool hands the next state vector
# of the same size the same blocks, whose handles then hit the kept mappings.  The other half of the
# contract is in the engine: a block that was exported is never freed behind the peers' backs
# (engine.cu MemPool::exported).  When a sharded state vector of ANOTHER size is created, every rank
# (they all take this decision alike) closes its kept mappings, the ranks meet, and only then the
# exported blocks are freed (qgb_pool_trim_exported) and the new shards allocated.
_RETAINED = {}      # id(library) -> {'bases': {mapping base: True}, 'key': shard shape they serve}


class DistContext:
    def __init__(self, local_module, group=None, exchange='auto', shard_min_lanes=None, push=True):
        if not dist.is_initialized():
            raise RuntimeError('torch.distributed is not initialized.')
        self.local = local_module
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.g = int(round(math.log2(self.world)))
        if (1 << self.g) != self.world:
            raise RuntimeError('world size must be a power of two, {}.'.format(self.world))
        self.api = local_module.api if hasattr(local_module, 'api') else local_module.get_api()
        self.on_cuda = self.api.backend_name.startswith('cuda')
        self.device = torch.device('cuda', torch.cuda.current_device()) if self.on_cuda \
            else torch.device('cpu')
        # The control plane (scalars, handles, barriers) follows the process group's backend: NCCL
        # works on device tensors, in stream order; gloo on host tensors.  gloo + the CUDA engine is
        # the "several ranks on ONE GPU" mode: NCCL refuses two ranks per device, CUDA IPC does not,
        # so the p2p exchange kernel still moves the amplitudes (tests/test_dist_shared_gpu.py; the
        # reference's logical devices on one GPU, tests/test_multidevice.py:13-21).
        self.comm_cuda = self.on_cuda and dist.get_backend(group) == 'nccl'
        self.comm_device = self.device if self.comm_cuda else torch.device('cpu')
        if exchange == 'auto':
            exchange = 'p2p' if self.on_cuda else 'collective'
        if exchange not in ('p2p', 'collective'):
            raise ValueError('exchange must be p2p or collective.')
        if exchange == 'collective' and self.on_cuda and not self.comm_cuda:
            raise ValueError('the collective exchange engine on GPUs needs the NCCL backend.')
        self.exchange = exchange
        # p2p engine: ranks PUSH into each other's spare buffers (write-only NVLink traffic) when
        # every rank can afford a spare buffer, else they swap in place (dist.cu)
        self.push = bool(push)
        if shard_min_lanes is None:
            shard_min_lanes = 24 if self.on_cuda else 12
        self.shard_min_lanes = max(int(shard_min_lanes), 2 * self.g + 2)
        self.stream = None
        self.stats = {'exchanges': 0, 'exchange_bytes': 0, 'exchange_lanes': 0,
                      'local_swaps': 0, 'sharded_joins': 0, 'gathered_operands': 0}
        self.timing = False          # bench.py: CUDA events around every exchange
        self._exchange_events = []
        self._barrier_buf = None
        self._keep = _RETAINED.setdefault(id(self.api.lib), {'bases': {}, 'key': None})
        # kept mappings need the engine to pin exported blocks; both only in the one-rank-per-GPU NCCL
        # configuration (ranks sharing a GPU over gloo are a test vehicle: they unmap at once, as before)
        self.keep_mappings = self.comm_cuda
        if self.keep_mappings:
            self.api.set_option('pin_exported', 1)

    def drop_retained(self):
        for base in list(self._keep['bases']):
            try:
                self.api.call('qgb_ipc_close', base)
            except Exception:
                pass
        self._keep['bases'].clear()
        self._keep['key'] = None

    # -- stream plumbing: NCCL work and the engine's kernels share one stream ----------------
    def bind_stream(self):
        if self.on_cuda and self.stream is None:
            if hasattr(self.local, 'module_init') and not getattr(self.local, 'initialized', True):
                self.local.module_init()           # the runtime module initialises lazily
            self.stream = torch.cuda.Stream(device=self.device)
            self.api.set_stream(self.stream.cuda_stream)

    def stream_ctx(self):
        if self.on_cuda:
            self.bind_stream()
            return torch.cuda.stream(self.stream)
        return _NullCtx()

    def device_barrier(self):
        """Orders the ranks ON THE STREAM (no host sync on GPUs): a 1-element all_reduce."""
        if self.on_cuda and not self.comm_cuda:
            torch.cuda.synchronize()                   # host-side control plane: drain, then meet
            dist.barrier(group=self.group)
            return
        with self.stream_ctx():
            if self._barrier_buf is None:
                self._barrier_buf = torch.zeros(1, dtype=torch.float32, device=self.comm_device)
            dist.all_reduce(self._barrier_buf, group=self.group)

    def all_reduce_sum(self, value):
        with self.stream_ctx():
            t = torch.tensor([value], dtype=torch.float64, device=self.comm_device)
            dist.all_reduce(t, group=self.group)
            return float(t.item())

    def all_gather_floats(self, value):
        with self.stream_ctx():
            t = torch.tensor([value], dtype=torch.float64, device=self.comm_device)
            out = torch.empty(self.world, dtype=torch.float64, device=self.comm_device)
            dist.all_gather_into_tensor(out, t, group=self.group)
            return out.cpu().numpy()

    def all_reduce_array(self, arr):
        """Sum of a NumPy array over the ranks (in place, returns it)."""
        flat = arr.view(np.float64) if arr.dtype == np.complex128 else \
            arr.view(np.float32) if arr.dtype == np.complex64 else arr
        with self.stream_ctx():
            t = torch.from_numpy(flat).to(self.comm_device)
            dist.all_reduce(t, group=self.group)
            flat[...] = t.cpu().numpy()
        return arr

    def all_gather_shards(self, whole, part):
        """whole = concatenation over the ranks of `part` (tensors over engine memory)."""
        with self.stream_ctx():
            if self.on_cuda and not self.comm_cuda:    # host-side control plane: stage through it
                self.stream.synchronize()
                gathered = torch.empty(whole.numel(), dtype=whole.dtype)
                dist.all_gather_into_tensor(gathered, part.cpu(), group=self.group)
                whole.copy_(gathered)
            else:
                dist.all_gather_into_tensor(whole, part, group=self.group)
            if self.on_cuda:
                self.stream.synchronize()

    def broadcast_float(self, value):
        with self.stream_ctx():
            t = torch.tensor([value], dtype=torch.float64, device=self.comm_device)
            dist.broadcast(t, src=dist.get_global_rank(self.group, 0) if self.group else 0,
                           group=self.group)
            return float(t.item())

    def broadcast_array(self, arr):
        with self.stream_ctx():
            t = torch.from_numpy(np.ascontiguousarray(arr)).to(self.comm_device)
            dist.broadcast(t, src=dist.get_global_rank(self.group, 0) if self.group else 0,
                           group=self.group)
            return t.cpu().numpy()


    def exchange_ms(self):
        """Device time of the exchanges recorded since the last call (needs timing = True)."""
        if self.on_cuda:
            torch.cuda.synchronize()
        total = sum(a.elapsed_time(b) for a, b in self._exchange_events)
        self._exchange_events = []
        return total


class _NullCtx:
    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


class DistQubitStates:
    """Protocol of NativeQubitStates (native_qubit_states.py:4-33) over a shard or a replica."""

    def __init__(self, ctx, dtype, local):
        self.ctx = ctx
        self.dtype = dtype
        self.local = local                   # NativeQubitStates of the local library
        self.processor = DistQubitProcessor(ctx, self)
        self.n_lanes = -1
        self.g = 0
        self.perm = []                       # logical lane -> physical lane
        self.pending = []
        self.peers = None                    # p2p: mapped pointers of all ranks' shards
        self.peers_alt = None                # ... and of their spare buffers (push exchange), or None
        self.peer_bases = []                 # the IPC mappings behind them (closed on delete)
        self.lane_states = []

    @property
    def ptr(self):
        return self.local.ptr

    @property
    def n_local(self):
        return self.n_lanes - self.g

    def rank_bit(self, physical):
        return (self.ctx.rank >> (physical - self.n_local)) & 1

    def delete(self):
        self._close_peers()
        if self.local is not None:
            self.local.delete()
            self.local = None

    def __del__(self):
        try:
            self.delete()
        except Exception:
            pass

    def _close_peers(self):
        if self.peers:
            keep = self.ctx._keep
            mine = getattr(self, 'peer_key', None)
            for base in self.peer_bases:
                if self.ctx.keep_mappings and mine is not None and mine == keep['key'] and base not in keep['bases']:
                    keep['bases'][base] = True      # the reference moves to the process-wide table
                    continue
                try:
                    self.ctx.api.call('qgb_ipc_close', base)
                except Exception:
                    pass
        self.peers = None
        self.peers_alt = None
        self.peer_bases = []

    def get_n_lanes(self):
        return self.n_lanes

    def reset_lane_states(self):
        self.lane_states = [-1] * self.n_lanes

    def get_lane_state(self, lane):
        return self.lane_states[lane]

    def set_lane_state(self, lane, value):
        self.lane_states[lane] = value

    def calc_probability(self, lane):
        return self.processor.calc_probability(self, lane)

    # -- raw memory ----------------------------------------------------------------------
    def data_ptr(self):
        ptr, nbytes = C.c_uint64(0), C.c_int64(0)
        self.ctx.api.call('qgb_qstates_data_ptr', self.local.ptr, C.byref(ptr), C.byref(nbytes))
        return ptr.value, nbytes.value

    def tensor(self):
        ptr, nbytes = self.data_ptr()
        return _tensor_view(ptr, nbytes, self.ctx.on_cuda, self.ctx.device)

    def alt_tensor(self):
        _, nbytes = self.data_ptr()
        ptr = C.c_uint64(0)
        self.ctx.api.call('qgb_qstates_alt_buffer', self.local.ptr, C.byref(ptr))
        return _tensor_view(ptr.value, nbytes, self.ctx.on_cuda, self.ctx.device)


class DistQubitProcessor:
    """Protocol of NativeQubitProcessor (native_qubit_processor.py:6-59)."""

    def __init__(self, ctx, qstates):
        self.ctx = ctx
        self.api = ctx.api
        self._qs = weakref.ref(qstates)

    # local processor of a qstates
    @staticmethod
    def _lp(qs):
        return qs.local.processor

    def delete(self):
        pass

    def reset(self):
        pass

    def synchronize(self):
        qs = self._qs()
        if qs is not None and qs.local is not None:
            self.flush(qs)
            self._lp(qs).synchronize()

    # -- allocation ---------------------------------------------------------------------------
    def initialize_qubit_states(self, qs, n_lanes):
        ctx = self.ctx
        qs._close_peers()
        qs.n_lanes = n_lanes
        qs.g = ctx.g if (ctx.world > 1 and n_lanes >= ctx.shard_min_lanes) else 0
        qs.perm = list(range(n_lanes))
        qs.pending = []
        ctx.bind_stream()
        if qs.g and ctx.keep_mappings:
            key = (n_lanes - qs.g, np.dtype(qs.dtype).itemsize)
            keep = ctx._keep
            if keep['key'] is not None and keep['key'] != key:
                # shards of another size: the kept mappings go, on every rank, and only when all of
                # them are gone are the blocks behind them freed (see _RETAINED)
                ctx.drop_retained()
                ctx.device_barrier()
                torch.cuda.synchronize()         # (the barrier is stream-ordered; freeing is a host call)
                self.api.call('qgb_pool_trim_exported')
            keep['key'] = key
            qs.peer_key = key
        self._lp(qs).initialize_qubit_states(qs.local, n_lanes - qs.g)
        qs.reset_lane_states()

    def reset_qubit_states(self, qs):
        qs.pending = []
        qs.perm = list(range(qs.n_lanes))
        self._lp(qs).reset_qubit_states(qs.local)
        if qs.g and self.ctx.rank != 0:
            self._scale_all(qs, 0., 0.)
        qs.reset_lane_states()

    def _scale_all(self, qs, re, im, local_ctrls=()):
        """Multiply every local amplitude (with the local controls set) by re + i im."""
        if local_ctrls:
            # one control becomes the target of diag(1, factor): a (controlled) phase gate, which the
            # engine folds into a neighbouring gate on that lane or into a phase fan
            ctrls = list(local_ctrls)
            lane = ctrls.pop()
            mat = (C.c_double * 8)(1., 0., 0., 0., 0., 0., re, im)
            self._apply_raw(qs, mat, ctrls, lane)
        else:
            mat = (C.c_double * 8)(re, im, 0., 0., 0., 0., re, im)
            self._apply_raw(qs, mat, [], 0)

    def _apply_raw(self, qs, mat, ctrls, target):
        lp = self._lp(qs)
        if ctrls:
            n = len(ctrls)
            rc = self.api.lib.qgb_qproc_apply_controlled_gate(lp.ptr, mat, qs.local.ptr,
                                                              (C.c_int * n)(*ctrls), n, target)
        else:
            rc = self.api.lib.qgb_qproc_apply_gate(lp.ptr, mat, qs.local.ptr, target)
        if rc:
            self.api.check(rc)

    # -- gates: deferred in logical coordinates -----------------------------------------------
    def _matrix(self, gate_type, adjoint):
        gate_id, cargs, n_args = NativeQubitProcessor._gate_key(gate_type)
        mat = (C.c_double * 8)()
        rc = self.api.lib.qgb_gate_matrix(gate_id, cargs, n_args, 1 if adjoint else 0, mat)
        if rc:
            self.api.check(rc)
        return mat

    # Streaming: every STREAM submitted gates, whatever needs no exchange goes on to the local engine
    # (which launches passes as its own queue fills), so the device works while the front end is still
    # dispatching; the gates that wait for an exchange stay here, in order.
    STREAM = int(os.environ.get('QGB_DIST_STREAM', '2048'))

    def _submitted(self, qs):
        qs.n_fresh = getattr(qs, 'n_fresh', 0) + 1
        if qs.n_fresh >= self.STREAM and qs.g:
            qs.n_fresh = 0
            qs.pending = self._apply_unblocked(qs, qs.pending)

    def apply_gate(self, gate_type, adjoint, qs, lane):
        qs.pending.append(Gate(self._matrix(gate_type, adjoint), (), lane))
        self._submitted(qs)

    def apply_controlled_gate(self, gate_type, adjoint, qs, ctrl_lanes, target_lane):
        qs.pending.append(Gate(self._matrix(gate_type, adjoint), tuple(ctrl_lanes), target_lane))
        self._submitted(qs)

    def submit_tuples(self, qs, gates):
        """bench.py's compact form: (u3 angles or None, control lane or -1, target lane)."""
        lib = self.api.lib
        uid, xid = _capi.GATE_IDS['U'], _capi.GATE_IDS['X']
        noargs = (C.c_double * 1)()
        xmat = (C.c_double * 8)()
        self.api.check(lib.qgb_gate_matrix(xid, noargs, 0, 0, xmat))
        for angles, ctrl, target in gates:
            if ctrl < 0:
                mat = (C.c_double * 8)()
                self.api.check(lib.qgb_gate_matrix(uid, (C.c_double * 3)(*angles), 3, 0, mat))
                qs.pending.append(Gate(mat, (), target))
            else:
                qs.pending.append(Gate(xmat, (ctrl,), target))

    def flush(self, qs):
        """Run the deferred gates: translate to physical lanes, exchange lanes where a
        non-diagonal gate targets a global lane, queue everything in the local engine and
        push it to the device."""
        self._run_pending(qs)
        lp = self._lp(qs)
        if hasattr(lp, 'flush'):
            lp.flush(qs.local)

    def _run_pending(self, qs):
        gates, qs.pending = qs.pending, []
        if not gates:
            return
        if qs.g == 0:
            for gate in gates:
                self._apply_raw(qs, gate.mat, [qs.perm[c] for c in gate.ctrls],
                                qs.perm[gate.target])
            return
        # Gates run as early as commutation allows: a non-diagonal gate on a global lane (and
        # everything behind it that does not commute with it) waits, the rest of the queue goes
        # on.  Only when nothing more can run do lanes trade places, so one exchange serves a
        # whole light cone of gates instead of one gate.
        pending = gates
        while pending:
            pending = self._apply_unblocked(qs, pending)
            if pending:
                self._swap_in(qs, pending)

    def _apply_unblocked(self, qs, pending):
        """Apply, in order, every gate that needs no exchange and commutes with all the gates
        skipped before it; returns the skipped gates (order kept)."""
        n_local = qs.n_local
        full = (1 << qs.n_lanes) - 1
        perm = qs.perm
        blocked_x = blocked_z = 0
        rest = []
        for idx, gate in enumerate(pending):
            if blocked_x == full:
                rest.extend(pending[idx:])
                break
            x, z = gate.xmask, gate.zmask
            if (x & (blocked_x | blocked_z)) or (z & blocked_x) or (x and perm[gate.target] >= n_local):
                rest.append(gate)
                blocked_x |= x
                blocked_z |= z
            else:
                self._apply_local(qs, gate)
        return rest

    def _apply_local(self, qs, gate):
        n_local = qs.n_local
        local_ctrls = []
        for c in gate.ctrls:
            p = qs.perm[c]
            if p >= n_local:
                if not qs.rank_bit(p):
                    return                      # a global control is 0 on this rank
            else:
                local_ctrls.append(p)
        tp = qs.perm[gate.target]
        if tp < n_local:
            self._apply_raw(qs, gate.mat, local_ctrls, tp)
            return
        # diagonal gate on a global lane: this rank's diagonal entry, as a phase
        m = gate.mat
        re, im = (m[6], m[7]) if qs.rank_bit(tp) else (m[0], m[1])
        if re == 1. and im == 0.:
            return
        self._scale_all(qs, re, im, local_ctrls)

    # -- lane exchange -------------------------------------------------------------------------
    LOOKAHEAD = 2048        # gates of the blocked queue the victim choice looks at

    def _swap_in(self, qs, pending):
        """Nothing in `pending` can run: trade every global lane a waiting non-diagonal gate
        targets for a local lane.  The victims are the local lanes whose absence blocks the
        least of the waiting queue (each candidate is scored by how many of the next LOOKAHEAD
        gates could still run if that lane alone were global)."""
        n_local = qs.n_local
        phys_to_logical = [0] * qs.n_lanes
        for logical, p in enumerate(qs.perm):
            phys_to_logical[p] = logical
        window = pending[:self.LOOKAHEAD]
        wanted = []                                   # global physical lanes, soonest use first
        for gate in window:
            p = qs.perm[gate.target]
            if gate.xmask and p >= n_local and p not in wanted:
                wanted.append(p)
        if not wanted:
            raise RuntimeError('blocked gate queue without a global target.')
        min_victim = 1 if np.dtype(qs.dtype) == np.float32 else 0   # 16-byte exchange units
        if n_local >= 12:
            min_victim = 5       # victims at lane >= 5 move runs of >= 512 bytes over the link
        cand_phys = list(range(min_victim, n_local))
        if not cand_phys:
            raise RuntimeError('no local lane available for an exchange.')
        cand = np.array([1 << phys_to_logical[p] for p in cand_phys], dtype=np.uint64)
        bx = np.zeros(len(cand), np.uint64)
        bz = np.zeros(len(cand), np.uint64)
        score = np.zeros(len(cand), np.int64)
        zero = np.uint64(0)
        for gate in window:
            x, z = np.uint64(gate.xmask), np.uint64(gate.zmask)
            blocked = ((x & (bx | bz)) != zero) | ((z & bx) != zero) | ((x & cand) != zero)
            score += ~blocked
            bx |= np.where(blocked, x, zero)
            bz |= np.where(blocked, z, zero)
        # best score first; among lanes within 10% of the best, the highest (the exchange moves
        # runs of 2^lane amplitudes: longer runs use the links better)
        best = int(score.max())
        order = sorted(range(len(cand_phys)),
                       key=lambda i: (-(best if int(score[i]) * 10 >= best * 9 else int(score[i])),
                                      -cand_phys[i]))
        pairs = [(g_pos, cand_phys[i]) for g_pos, i in zip(wanted, order)]
        self.exchange(qs, pairs)

    def make_local(self, qs, lane):
        """Force one logical lane onto a local physical lane (measurement collapse, reset)."""
        self._run_pending(qs)
        p = qs.perm[lane]
        if p < qs.n_local:
            return
        self.exchange(qs, [(p, qs.n_local - 1)])

    def exchange(self, qs, pairs):
        """pairs: [(global physical lane, local physical lane)]; afterwards the logical lanes
        that sat on them have traded places."""
        ctx = self.ctx
        n_local = qs.n_local
        k = len(pairs)
        phys_to_logical = [0] * qs.n_lanes
        for logical, p in enumerate(qs.perm):
            phys_to_logical[p] = logical
        if ctx.exchange == 'collective':
            # the collective engine moves whole contiguous blocks: victims must be the top k
            # local lanes.  A victim elsewhere first changes places with a top lane inside the
            # shard: three CX gates, fused by the engine into one pass.
            tops = list(range(n_local - k, n_local))
            chosen = [l for _, l in pairs]
            free_tops = [t for t in tops if t not in chosen]
            fixed = []
            for g_pos, l_pos in pairs:
                if l_pos in tops:
                    fixed.append((g_pos, l_pos))
                    continue
                top = free_tops.pop()
                self._local_swap(qs, l_pos, top, phys_to_logical)
                fixed.append((g_pos, top))
            pairs = fixed
        pairs = sorted(pairs, key=lambda gl: gl[1])          # ascending victim lane
        sel_bits = [g_pos - n_local for g_pos, _ in pairs]     # rank bit of selector bit i
        victims = [l_pos for _, l_pos in pairs]
        my_sel = 0
        for i, b in enumerate(sel_bits):
            my_sel |= ((ctx.rank >> b) & 1) << i
        base_rank = ctx.rank
        for b in sel_bits:
            base_rank &= ~(1 << b)

        def rank_of(sel):
            r = base_rank
            for i, b in enumerate(sel_bits):
                r |= ((sel >> i) & 1) << b
            return r

        # everything queued so far must be in the shard before it moves
        lp = self._lp(qs)
        if hasattr(lp, 'flush'):
            lp.flush(qs.local)
        _, nbytes = qs.data_ptr()
        ev0 = ev1 = None
        if ctx.timing and ctx.on_cuda:
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record(ctx.stream)
        if ctx.exchange == 'p2p':
            self._ensure_peers(qs)
            table = qs.peers_alt if qs.peers_alt is not None else qs.peers
            peer_ptrs = (C.c_uint64 * (1 << k))(*[table[rank_of(sel)] for sel in range(1 << k)])
            ctx.device_barrier()                 # all ranks' passes are ahead of the kernel
            with ctx.stream_ctx():
                if qs.peers_alt is not None:     # push into the spare buffers, then everybody flips
                    self.api.call('qgb_qstates_exchange_push', qs.local.ptr, peer_ptrs, k,
                                  (C.c_int * k)(*victims), my_sel)
                    qs.peers, qs.peers_alt = qs.peers_alt, qs.peers
                    ctx.stats['push_exchanges'] = ctx.stats.get('push_exchanges', 0) + 1
                else:
                    self.api.call('qgb_qstates_exchange_p2p', qs.local.ptr, peer_ptrs, k,
                                  (C.c_int * k)(*victims), my_sel)
            ctx.device_barrier()                 # nobody touches its shard while a peer still does
        else:
            src = qs.tensor()
            dst = qs.alt_tensor()
            block = src.numel() >> k
            ops, keep = [], None
            for sel in range(1 << k):
                sl = slice(sel * block, (sel + 1) * block)
                if sel == my_sel:
                    keep = sl
                    continue
                peer = rank_of(sel)
                peer_global = dist.get_global_rank(ctx.group, peer) if ctx.group else peer
                ops.append(dist.P2POp(dist.isend, src[sl], peer_global, group=ctx.group))
                ops.append(dist.P2POp(dist.irecv, dst[sl], peer_global, group=ctx.group))
            with ctx.stream_ctx():
                reqs = dist.batch_isend_irecv(ops)
                dst[keep].copy_(src[keep])
                for req in reqs:
                    req.wait()
            self.api.call('qgb_qstates_flip', qs.local.ptr)
        if ev0 is not None:
            ev1.record(ctx.stream)
            ctx._exchange_events.append((ev0, ev1))
        # bookkeeping: the logical lanes trade physical places
        for g_pos, l_pos in pairs:
            lg, ll = phys_to_logical[g_pos], phys_to_logical[l_pos]
            qs.perm[lg], qs.perm[ll] = l_pos, g_pos
            phys_to_logical[g_pos], phys_to_logical[l_pos] = ll, lg
        ctx.stats['exchanges'] += 1
        ctx.stats['exchange_lanes'] += k
        ctx.stats['exchange_bytes'] += nbytes - (nbytes >> k)

    def _local_swap(self, qs, a, b, phys_to_logical):
        x = (C.c_double * 8)(0., 0., 1., 0., 1., 0., 0., 0.)
        self._apply_raw(qs, x, [a], b)
        self._apply_raw(qs, x, [b], a)
        self._apply_raw(qs, x, [a], b)
        la, lb = phys_to_logical[a], phys_to_logical[b]
        qs.perm[la], qs.perm[lb] = b, a
        phys_to_logical[a], phys_to_logical[b] = lb, la
        self.ctx.stats['local_swaps'] += 1

    def _ensure_peers(self, qs):
        if qs.peers is not None:
            return
        ctx = self.ctx
        qs.peers, bases = self._map_peers(qs, 'qgb_qstates_ipc_export')
        qs.peer_bases = bases
        qs.peers_alt = None
        if ctx.push:
            # the push exchange needs a spare buffer on EVERY rank (twice the shard): ask, agree
            handle = (C.c_ubyte * 64)()
            offset = C.c_int64(0)
            _, shard_bytes = qs.data_ptr()
            fits = True
            if ctx.on_cuda:                      # (a doomed 128 GiB allocation is not worth attempting)
                fits = 2 * shard_bytes <= 0.95 * torch.cuda.mem_get_info()[1]
            try:
                if not fits:
                    raise RuntimeError('no room for a spare buffer')
                self.api.call('qgb_qstates_ipc_export_alt', qs.local.ptr, handle, C.byref(offset))
                mine = 1.
            except RuntimeError:                 # out of device memory: swap in place instead
                mine = 0.
            everyone = ctx.all_gather_floats(mine)
            if everyone.min() > 0.:
                qs.peers_alt, alt_bases = self._map_peers(qs, 'qgb_qstates_ipc_export_alt')
                qs.peer_bases = qs.peer_bases + alt_bases

    def _map_peers(self, qs, export_fn):
        """IPC-map one buffer (the shard or its spare) of every rank; returns ([pointer per rank],
        [mappings to close])."""
        ctx = self.ctx
        handle = (C.c_ubyte * 64)()
        offset = C.c_int64(0)
        self.api.call(export_fn, qs.local.ptr, handle, C.byref(offset))
        # everything on the engine's stream, the host read included: a read issued on another
        # stream would race with the all-gather and open a garbage handle
        with ctx.stream_ctx():
            mine = torch.tensor(list(bytes(handle)) + [offset.value], dtype=torch.int64,
                                device=ctx.comm_device)
            everyone = torch.empty(ctx.world * 65, dtype=torch.int64, device=ctx.comm_device)
            dist.all_gather_into_tensor(everyone, mine, group=ctx.group)
            table = everyone.cpu().numpy().reshape(ctx.world, 65)
        if export_fn.endswith('_alt'):
            own = C.c_uint64(0)
            self.api.call('qgb_qstates_alt_buffer', qs.local.ptr, C.byref(own))
            own_ptr = own.value
        else:
            own_ptr, _ = qs.data_ptr()
        peers, bases = [], []
        for r in range(ctx.world):
            if r == ctx.rank:
                peers.append(own_ptr)
                continue
            raw = (C.c_ubyte * 64)(*[int(v) for v in table[r, :64]])
            base = C.c_uint64(0)
            self.api.call('qgb_ipc_open', raw, C.byref(base))
            bases.append(base.value)
            peers.append(base.value + int(table[r, 64]))
        return peers, bases

    # -- observers ----------------------------------------------------------------------------
    def calc_probability(self, qs, lane):
        self._run_pending(qs)
        lp = self._lp(qs)
        if qs.g == 0:
            return lp.calc_probability(qs.local, qs.perm[lane])
        p = qs.perm[lane]
        if p < qs.n_local:
            part = lp.calc_probability(qs.local, p)
        else:
            norm = C.c_double(0.)
            self.api.call('qgb_qproc_calc_norm', lp.ptr, qs.local.ptr, C.byref(norm))
            part = 0. if qs.rank_bit(p) else norm.value
        return self.ctx.all_reduce_sum(part)

    def decohere(self, value, prob, qs, lane):
        self._run_pending(qs)
        p = qs.perm[lane]
        if p < qs.n_local:
            self._lp(qs).decohere(value, prob, qs.local, p)
            return
        # the measured lane is a rank bit: ranks on the other side go to zero, the others rescale
        # (factor computed in double, like CPUQubitProcessor.cpp:227,235)
        if qs.rank_bit(p) == int(value):
            norm = 1. / math.sqrt(prob) if value == 0 else 1. / math.sqrt(1. - prob)
            self._scale_all(qs, norm, 0.)
        else:
            self._scale_all(qs, 0., 0.)

    def apply_reset(self, qs, lane):
        self.make_local(qs, lane)
        self._lp(qs).apply_reset(qs.local, qs.perm[lane])

    def decohere_and_separate(self, value, prob, qs0, qs1, qs, lane):
        self.make_local(qs, lane)
        p = qs.perm[lane]
        lp = self._lp(qs)
        # lane map of the remainder: `lane` leaves, every lane above it moves down by one
        rest_perm = []
        for logical in range(qs.n_lanes):
            if logical == lane:
                continue
            q = qs.perm[logical]
            rest_perm.append(q - 1 if q > p else q)
        if qs.g == 0 or qs0.g == qs.g:
            lp.decohere_and_separate(value, prob, qs0.local, qs1.local, qs.local, p)
        else:
            # the remainder falls below the sharding threshold: separate into a temporary shard,
            # then gather the shards into the replica (rank-major = global lanes on top)
            tmp = self.ctx.local.create_qubit_states(qs.dtype)
            tmp.processor.initialize_qubit_states(tmp, qs.n_local - 1)
            lp.decohere_and_separate(value, prob, tmp, qs1.local, qs.local, p)
            ptr, nbytes = C.c_uint64(0), C.c_int64(0)
            self.api.call('qgb_qstates_data_ptr', tmp.ptr, C.byref(ptr), C.byref(nbytes))
            part = _tensor_view(ptr.value, nbytes.value, self.ctx.on_cuda, self.ctx.device)
            self.ctx.all_gather_shards(qs0.tensor(), part)
            tmp.delete()
        qs0.perm = rest_perm
        qs0.pending = []
        qs1.perm = [0]
        qs1.pending = []

    def _gather_replica(self, src):
        """A sharded qstates as a local, unsharded copy on every rank (all-gather of the shards:
        rank-major = the global lanes on top, so the physical lane order is unchanged)."""
        ctx = self.ctx
        tmp = ctx.local.create_qubit_states(src.dtype)
        tmp.processor.initialize_qubit_states(tmp, src.n_lanes)
        ptr, nbytes = C.c_uint64(0), C.c_int64(0)
        self.api.call('qgb_qstates_data_ptr', tmp.ptr, C.byref(ptr), C.byref(nbytes))
        whole = _tensor_view(ptr.value, nbytes.value, ctx.on_cuda, ctx.device)
        ctx.all_gather_shards(whole, src.tensor())
        return tmp

    def join(self, qs, qs_list, n_new):
        """Kronecker product of the groups (qubits_handler.py:60-89; the LAST list element holds
        the lowest lanes).  Replicated operands multiply locally.  One sharded operand keeps its
        global lanes: every rank multiplies ITS shard with the replicated operands, no traffic.
        Further sharded operands are first gathered into replicas (they must fit one GPU)."""
        for src in qs_list:
            self._run_pending(src)
        lp = self._lp(qs)
        sharded = [src for src in qs_list if src.g]
        temps = []
        if not sharded:
            srcs = [src.local for src in qs_list]
            if qs.g == 0:
                lp.join(qs.local, srcs, n_new)
            else:
                ptrs = self.api.handle_array([s.ptr for s in srcs])
                self.api.call('qgb_qproc_join_shard', lp.ptr, qs.local.ptr, ptrs, len(srcs), n_new,
                              qs.n_lanes, self.ctx.rank << qs.n_local)
            keep = None
        else:
            if qs.g == 0:
                raise RuntimeError('join: a sharded operand needs a sharded destination.')
            keep = max(sharded, key=lambda src: src.n_lanes)
            srcs = []
            for src in qs_list:
                if src is keep or not src.g:
                    srcs.append(src.local)
                else:
                    rep = self._gather_replica(src)
                    temps.append(rep)
                    srcs.append(rep)
            lp.join(qs.local, srcs, n_new)
            self.ctx.stats['sharded_joins'] += 1
            self.ctx.stats['gathered_operands'] += len(temps)
        # logical lane -> physical lane of the product
        perm = [0] * qs.n_lanes
        off_log = off_phys = 0
        for src in reversed(qs_list):
            if src is keep:
                for logical, p in enumerate(src.perm):
                    perm[off_log + logical] = off_phys + p if p < src.n_local \
                        else qs.n_local + (p - src.n_local)
                off_phys += src.n_local
            else:
                for logical, p in enumerate(src.perm):
                    perm[off_log + logical] = off_phys + p
                off_phys += src.n_lanes
            off_log += src.n_lanes
        for i in range(qs.n_lanes - off_log):          # new |0> qubits: the top LOCAL lanes
            perm[off_log + i] = off_phys + i
        qs.perm = perm
        qs.pending = []
        for rep in temps:
            rep.delete()


class DistSamplingPool:
    """Pool over a sharded probability vector: the rank that owns a draw searches it."""

    def __init__(self, ctx, local_pool, qreg_ordering, empty_lanes, n_pool_lanes, ends, fp32):
        self.ctx = ctx
        self.local_pool = local_pool          # NativeSamplingPool over this rank's slice
        self.qreg_ordering = qreg_ordering
        self.empty_lanes = sorted(empty_lanes)
        self.n_pool_lanes = n_pool_lanes
        self.ends = ends                      # cumulative probability at the end of each rank
        self.fp32 = fp32
        self.mask = NativeQubitsStatesGetter.create_lane_mask(empty_lanes)

    def delete(self):
        if self.local_pool is not None:
            self.local_pool.delete()
            self.local_pool = None

    def sample(self, n_samples, randnum=None):
        ctx = self.ctx
        if randnum is None:
            randnum = np.random.random_sample([n_samples])
            randnum = ctx.broadcast_array(randnum)     # every rank must search the same draws
        randnum = np.ascontiguousarray(randnum, np.float64)
        if randnum.size < n_samples:
            raise ValueError('array size too small.')
        r = randnum[:n_samples]
        key = r.astype(np.float32).astype(np.float64) if self.fp32 else r
        owner = np.minimum(np.searchsorted(self.ends, key, side='right'), ctx.world - 1)
        mine = np.flatnonzero(owner == ctx.rank)
        obs = np.zeros([n_samples], np.int64)
        if mine.size:
            local = self.local_pool.sample(int(mine.size), np.ascontiguousarray(r[mine])).intarray
            n_local_bits = self.n_pool_lanes - ctx.g
            obs[mine] = local | (np.int64(ctx.rank) << np.int64(n_local_bits))
        with ctx.stream_ctx():
            t = torch.from_numpy(obs).to(ctx.comm_device)
            dist.all_reduce(t, group=ctx.group)
            obs = t.cpu().numpy()
        for pos in self.empty_lanes:                    # deposit a 0 at every empty lane
            low = obs & ((np.int64(1) << np.int64(pos)) - 1)
            obs = ((obs - low) << 1) | low
        return ObservationList(self.qreg_ordering, obs, self.mask)


class DistQubitsStatesGetter:
    """Protocol of NativeQubitsStatesGetter (native_qubits_states_getter.py:6-102)."""

    def __init__(self, ctx, dtype, local_getter):
        self.ctx = ctx
        self.dtype = dtype
        self.local = local_getter

    def delete(self):
        if self.local is not None:
            self.local.delete()
            self.local = None

    @staticmethod
    def _flush(lane_trans):
        for qs, _ in lane_trans:
            qs.processor._run_pending(qs)

    def get_states(self, values, offset, mathop, lane_trans, empty_lanes, n_states, start, step):
        self._flush(lane_trans)
        ctx = self.ctx
        if not any(qs.g for qs, _ in lane_trans):
            phys = [(qs.local, [_Lane(qs.perm[l.local], l.external) for l in lanes])
                    for qs, lanes in lane_trans]
            for _, lanes in phys:
                lanes.sort(key=lambda l: l.local)
            self.local.get_states(values, offset, mathop, phys, empty_lanes, n_states, start, step)
            return
        if values.size - offset < n_states:
            raise ValueError('array size too small.')
        n_qregs = sum(len(lanes) for _, lanes in lane_trans) + len(empty_lanes)
        if not (0 <= start < (1 << n_qregs)) or not (0 <= start + step * (n_states - 1) < (1 << n_qregs)):
            raise ValueError('value out of range')
        ext = start + step * np.arange(n_states, dtype=np.int64)
        out = None
        for qs, lanes in lane_trans:
            n_local = qs.n_local
            local_lanes = [_Lane(qs.perm[l.local], l.external) for l in lanes
                           if qs.perm[l.local] < n_local]
            local_lanes.sort(key=lambda l: l.local)
            factor = np.empty([n_states], values.dtype)
            # a one-qstates request over the same external index space
            self._single(factor, mathop, qs.local, local_lanes, empty_lanes, n_qregs, n_states,
                         start, step)
            if qs.g:
                owned = np.ones([n_states], bool)
                for l in lanes:
                    p = qs.perm[l.local]
                    if p >= n_local:
                        owned &= ((ext >> l.external) & 1) == qs.rank_bit(p)
                factor[~owned] = 0
                ctx.all_reduce_array(factor)
            out = factor if out is None else out * factor
        values[offset:offset + n_states] = out

    def _single(self, factor, mathop, local_qs, local_lanes, empty_lanes, n_qregs, n_states, start,
                step):
        api = self.ctx.api
        table = [0] * len(local_lanes)
        for l in local_lanes:
            table[l.local] = l.external
        mask = NativeQubitsStatesGetter.create_lane_mask(empty_lanes)
        api.call('qgb_getter_get_states', self.local.ptr, factor.ctypes.data_as(C.c_void_p), 0,
                 _mathop_id(mathop), api.int_array(table), api.int_array([len(table)]), int(mask),
                 api.handle_array([local_qs.ptr]), 1, n_qregs, int(n_states), int(start), int(step))

    # -- sampling pools -------------------------------------------------------------------------
    def create_sampling_pool(self, qreg_ordering, n_lanes, n_hidden_lanes, lane_trans, empty_lanes,
                             sampling_pool_factory=None):
        self._flush(lane_trans)
        ctx = self.ctx
        sharded = [qs for qs, _ in lane_trans if qs.g]
        if not sharded:
            phys = [(qs.local, sorted([_Lane(qs.perm[l.local], l.external) for l in lanes],
                                      key=lambda l: l.local)) for qs, lanes in lane_trans]
            return self.local.create_sampling_pool(qreg_ordering, n_lanes, n_hidden_lanes, phys,
                                                   empty_lanes, sampling_pool_factory)
        if len(lane_trans) == 1 and sampling_pool_factory is None and n_lanes >= 2 * ctx.g:
            return self._sharded_pool(qreg_ordering, n_lanes, n_hidden_lanes, lane_trans[0],
                                      empty_lanes)
        prob = self._replicated_prob(n_lanes, lane_trans)
        if sampling_pool_factory is not None:
            return sampling_pool_factory(prob.astype(self.dtype), empty_lanes, qreg_ordering)
        pool = C.c_uint64(0)
        ctx.api.call('qgb_pool_from_prob_array', _capi.prec_of(self.dtype),
                     prob.ctypes.data_as(C.POINTER(C.c_double)), n_lanes,
                     ctx.api.int_array(empty_lanes), len(empty_lanes), C.byref(pool))
        return NativeSamplingPool(ctx.api, pool.value, qreg_ordering,
                                  NativeQubitsStatesGetter.create_lane_mask(empty_lanes))

    def _sharded_pool(self, qreg_ordering, n_lanes, n_hidden, entry, empty_lanes):
        """One sharded qstates: make the pool's top g lanes the rank bits (in order), so that the
        ranks own consecutive slices of the probability vector in pool order."""
        ctx = self.ctx
        qs, lanes = entry
        g, n_local = qs.g, qs.n_local
        by_ext = {l.external: l.local for l in lanes if l.external >= 0}
        want = [by_ext[n_lanes - g + i] for i in range(g)]          # logical lane for rank bit i
        proc = qs.processor
        # phase 1: every wanted lane that sits on the wrong global position becomes local
        movers = [(qs.perm[lg], lg) for lg in want if qs.perm[lg] >= n_local
                  and qs.perm[lg] != n_local + want.index(lg)]
        if movers:
            used = set(qs.perm[lg] for lg in want)
            free = [p for p in range(n_local - 1, 0, -1) if p not in used]
            proc.exchange(qs, [(p, free.pop(0)) for p, _ in movers])
        # phase 2: bring each wanted lane to its rank bit
        pairs = [(n_local + i, qs.perm[lg]) for i, lg in enumerate(want)
                 if qs.perm[lg] != n_local + i]
        if pairs:
            proc.exchange(qs, pairs)
        # local slice: local pool lanes in pool order above the hidden lanes
        table = [0] * n_local
        hidden_idx = 0
        for l in sorted(lanes, key=lambda l: qs.perm[l.local]):
            p = qs.perm[l.local]
            if p >= n_local:
                continue
            if l.external < 0:
                table[p] = hidden_idx
                hidden_idx += 1
            else:
                table[p] = l.external + n_hidden
        api = ctx.api
        pool, total = C.c_uint64(0), C.c_double(0.)
        api.call('qgb_getter_create_sampling_pool_partial', self.local.ptr, api.int_array(table),
                 api.int_array([n_local]), api.handle_array([qs.local.ptr]), 1, n_lanes - g,
                 n_hidden, C.byref(pool), C.byref(total))
        totals = ctx.all_gather_floats(total.value)
        offsets = np.concatenate([[0.], np.cumsum(totals)[:-1]])
        grand = 0.
        for t in totals:                                  # rank order, left to right
            grand += float(t)
        if not abs(grand - 1.) <= 0.05:                # CPUSamplingPool.cpp:31-34
            api.call('qgb_pool_delete', pool.value)
            raise RuntimeError('error in probability sum is beyond 0.05., {}.'.format(grand))
        api.call('qgb_pool_finalize', pool.value, float(offsets[ctx.rank]), grand)
        ends = (offsets + totals) * (1. / grand)
        local_pool = NativeSamplingPool(api, pool.value, qreg_ordering, 0)
        return DistSamplingPool(ctx, local_pool, qreg_ordering, empty_lanes, n_lanes, ends,
                                np.dtype(self.dtype) == np.float32)

    def _replicated_prob(self, n_lanes, lane_trans):
        """Whole marginal probability vector on every rank (small pools, several qstates):
        P[ext] = prod over qstates of its own marginal over its pool lanes."""
        ctx, api = self.ctx, self.ctx.api
        idx = np.arange(1 << n_lanes, dtype=np.int64)
        prob = None
        for qs, lanes in lane_trans:
            n_local = qs.n_local
            pool_lanes = sorted([l for l in lanes if l.external >= 0], key=lambda l: l.external)
            hidden = [l for l in lanes if l.external < 0]
            order = {l.local: i for i, l in enumerate(pool_lanes)}     # logical -> rank in P_q
            local_pool = [l for l in pool_lanes if qs.perm[l.local] < n_local]
            local_hidden = [l for l in hidden if qs.perm[l.local] < n_local]
            table = [0] * n_local
            for i, l in enumerate(local_hidden):
                table[qs.perm[l.local]] = i
            for i, l in enumerate(local_pool):
                table[qs.perm[l.local]] = len(local_hidden) + i
            part = np.empty([1 << len(local_pool)], self.dtype)
            api.call('qgb_getter_prepare_prob_array', self.local.ptr,
                     part.ctypes.data_as(C.c_void_p), api.int_array(table),
                     api.int_array([n_local]), api.handle_array([qs.local.ptr]), 1,
                     len(local_pool), len(local_hidden))
            part = part.astype(np.float64)
            if qs.g:
                # scatter this rank's slice into the qstates' marginal, then sum over the ranks
                full = np.zeros([1 << len(pool_lanes)], np.float64)
                li = np.arange(1 << len(local_pool), dtype=np.int64)
                pos = np.zeros_like(li)
                for i, l in enumerate(local_pool):
                    pos |= ((li >> i) & 1) << order[l.local]
                for l in pool_lanes:
                    p = qs.perm[l.local]
                    if p >= n_local and qs.rank_bit(p):
                        pos |= np.int64(1) << order[l.local]
                np.add.at(full, pos, part)
                # ranks that differ only in hidden global lanes add up; ranks that differ in a
                # pool lane wrote disjoint entries
                ctx.all_reduce_array(full)
                part = full
            sub = np.zeros_like(idx)
            for i, l in enumerate(pool_lanes):
                sub |= ((idx >> l.external) & 1) << i
            factor = part[sub]
            prob = factor if prob is None else prob * factor
        return np.ascontiguousarray(prob, np.float64)


class DistRuntimeModule:
    """Module-level runtime protocol (cudaruntime.py:24-91) for sharded execution."""

    def __init__(self, local_module, **options):
        self.local = local_module
        self._options = options
        self._ctx = None

    @property
    def ctx(self):
        if self._ctx is None:
            self._ctx = DistContext(self.local, **self._options)
        return self._ctx

    @property
    def api(self):
        return self.local.api

    def create_qubit_states(self, dtype):
        return DistQubitStates(self.ctx, dtype, self.local.create_qubit_states(dtype))

    def create_qubits_states_getter(self, dtype):
        return DistQubitsStatesGetter(self.ctx, dtype,
                                      self.local.create_qubits_states_getter(dtype))

    def broadcast_random(self, value):
        return self.ctx.broadcast_float(value)


def runtime(local_module, **options):
    """Sharded runtime over a local runtime module (qgate_b200.cudaruntime on GPUs; the tests
    use the reference-CPU shim with the gloo backend)."""
    return DistRuntimeModule(local_module, **options)
