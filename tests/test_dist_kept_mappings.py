"""Host side of the kept CUDA-IPC mappings (qgate_b200/dist.py), on fakes: which
mappings a deleted sharded state vector hands to the process-wide table, which it closes, and the order
of events when a sharded state vector of another size is created (mappings closed -> ranks meet ->
exported blocks freed -> allocation).  The engine's half of the contract is tests/native/pool_emul.cpp;
the real thing runs on 2 / 4 GPUs (tests/test_dist_nccl.py, tools/dist_mem_sequence.py)."""
import numpy as np

from qgate_b200 import dist as D


class FakeApi:
    def __init__(self, log):
        self.log = log
        self.lib = object()

    def call(self, name, *args):
        self.log.append((name,) + tuple(a for a in args if isinstance(a, int)))

    def set_option(self, name, value):
        self.log.append(('set_option', name, value))


class FakeLocalProcessor:
    def __init__(self, log):
        self.log = log

    def initialize_qubit_states(self, local, n_lanes):
        self.log.append(('local_initialize', n_lanes))


class FakeLocal:
    def __init__(self, log):
        self.processor = FakeLocalProcessor(log)
        self.ptr = 1

    def delete(self):
        pass


def make_ctx(log, keep_mappings=True):
    ctx = object.__new__(D.DistContext)
    ctx.api = FakeApi(log)
    ctx.rank, ctx.world, ctx.g = 0, 2, 1
    ctx.on_cuda = False          # (no torch.cuda calls in this test)
    ctx.comm_cuda = False
    ctx.keep_mappings = keep_mappings
    ctx.shard_min_lanes = 6
    ctx.stream = None
    ctx._keep = {'bases': {}, 'key': None}
    ctx.bind_stream = lambda: None
    ctx.device_barrier = lambda: log.append(('barrier',))
    return ctx


def new_state(ctx, log, n_lanes, dtype=np.float64):
    qs = D.DistQubitStates(ctx, dtype, FakeLocal(log))
    qs.processor.api = ctx.api
    qs.processor.initialize_qubit_states(qs, n_lanes)
    qs.reset_lane_states = lambda: None
    return qs


def with_peers(qs, bases):
    qs.peers, qs.peers_alt, qs.peer_bases = [0, bases[0]], None, list(bases)


def test_mappings_move_to_the_table_and_are_reused(monkeypatch):
    monkeypatch.setattr(D.torch.cuda, 'synchronize', lambda: None)
    log = []
    ctx = make_ctx(log)
    a = new_state(ctx, log, 10)
    assert ctx._keep['key'] == (9, 8) and a.peer_key == (9, 8)
    with_peers(a, [0x1000, 0x2000])
    a._close_peers()
    assert ctx._keep['bases'] == {0x1000: True, 0x2000: True}
    assert not [e for e in log if e[0] == 'qgb_ipc_close']          # nothing was unmapped
    # the next state vector of the same size opens the same handles: the engine's table returns the
    # same bases with one more reference each, which this state gives back when it goes
    b = new_state(ctx, log, 10)
    with_peers(b, [0x1000, 0x2000])
    b._close_peers()
    assert [e for e in log if e[0] == 'qgb_ipc_close'] == [('qgb_ipc_close', 0x1000), ('qgb_ipc_close', 0x2000)]
    assert ctx._keep['bases'] == {0x1000: True, 0x2000: True}
    assert not [e for e in log if e[0] in ('barrier', 'qgb_pool_trim_exported')]


def test_another_shard_size_drops_everything_in_order(monkeypatch):
    monkeypatch.setattr(D.torch.cuda, 'synchronize', lambda: None)
    log = []
    ctx = make_ctx(log)
    a = new_state(ctx, log, 10)
    with_peers(a, [0x1000, 0x2000])
    a._close_peers()
    del log[:]
    b = new_state(ctx, log, 12)                                          # other shard size
    names = [e[0] for e in log]
    assert names == ['qgb_ipc_close', 'qgb_ipc_close', 'barrier', 'qgb_pool_trim_exported', 'local_initialize'], names
    assert ctx._keep == {'bases': {}, 'key': (11, 8)} and b.peer_key == (11, 8)
    # same lanes, other precision: another size as well
    del log[:]
    new_state(ctx, log, 12, np.float32)
    assert [e[0] for e in log] == ['barrier', 'qgb_pool_trim_exported', 'local_initialize']
    # replicated (unsharded) state vectors do not take part
    del log[:]
    small = new_state(ctx, log, 4)
    assert small.g == 0 and [e[0] for e in log] == ['local_initialize'] and ctx._keep["key"] == (11, 4)


def test_without_kept_mappings_everything_is_closed_at_once(monkeypatch):
    monkeypatch.setattr(D.torch.cuda, 'synchronize', lambda: None)
    log = []
    ctx = make_ctx(log, keep_mappings=False)          # ranks sharing one GPU over gloo
    a = new_state(ctx, log, 10)
    with_peers(a, [0x1000, 0x2000])
    a._close_peers()
    assert [e for e in log if e[0] == 'qgb_ipc_close'] == [('qgb_ipc_close', 0x1000), ('qgb_ipc_close', 0x2000)]
    assert ctx._keep == {'bases': {}, 'key': None}
    del log[:]
    new_state(ctx, log, 12)
    assert [e[0] for e in log] == ['local_initialize']
