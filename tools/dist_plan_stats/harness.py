import sys, math, ctypes as C
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from qgate_b200 import dist as D, circuits, _capi
import qgate_b200.script as S
from qgate_b200 import model

class FakeCtx:
    def __init__(self, rank, size):
        self.rank, self.size = rank, size
        self.exchange = 'p2p'; self.stats = {'exchanges': 0, 'exchange_lanes': 0, 'exchange_bytes': 0, 'local_swaps': 0}
        self.timing = False; self.on_cuda = False

class Rec(D.DistQubitProcessor):
    def __init__(self, ctx, qs, out):
        self.ctx = ctx; self.out = out
    def _apply_raw(self, qs, mat, ctrls, target):
        cm = 0
        for c in ctrls: cm |= 1 << c
        self.out.write('G %d %d %s\n' % (target, cm, ' '.join(repr(float(mat[i])) for i in range(8))))
    def exchange(self, qs, pairs):
        self.out.write('F\n')
        phys_to_logical = [0] * qs.n_lanes
        for logical, p in enumerate(qs.perm): phys_to_logical[p] = logical
        for g_pos, l_pos in pairs:
            lg, ll = phys_to_logical[g_pos], phys_to_logical[l_pos]
            qs.perm[lg], qs.perm[ll] = l_pos, g_pos
            phys_to_logical[g_pos], phys_to_logical[l_pos] = ll, lg
        self.ctx.stats['exchanges'] += 1

class QS:
    pass

def gate_matrix(op):
    import qgate_b200.native as N
    return None

def main():
    n_gpus = int(sys.argv[1]); depth = int(sys.argv[2]); path = sys.argv[3]
    g = int(round(math.log2(n_gpus))); n = 30 + g
    rng = np.random.RandomState(1234)
    ctx = FakeCtx(0 if len(sys.argv) < 5 else int(sys.argv[4]), n_gpus)
    qs = QS(); qs.n_lanes = n; qs.g = g; qs.perm = list(range(n)); qs.pending = []; qs.dtype = np.float64
    qs.n_local = n - g
    qs.rank_bit = lambda physical: (ctx.rank >> (physical - qs.n_local)) & 1
    out = open(path, 'w')
    proc = Rec(ctx, qs, out)
    def u3(theta, phi, lam):
        c, s = math.cos(theta / 2), math.sin(theta / 2)
        m = [c, 0., -math.cos(lam) * s, -math.sin(lam) * s, math.cos(phi) * s, math.sin(phi) * s, math.cos(phi + lam) * c, math.sin(phi + lam) * c]
        return (C.c_double * 8)(*m)
    xmat = (C.c_double * 8)(0., 0., 1., 0., 1., 0., 0., 0.)
    for d in range(depth):
        for i in range(n):
            theta, phi, lam = rng.uniform(0., 2. * math.pi, 3)
            qs.pending.append(D.Gate(u3(theta, phi, lam), (), i))
        for i in range(d % 2, n - 1, 2):
            qs.pending.append(D.Gate(xmat, (i,), i + 1))
    proc._run_pending(qs)
    out.write('F\n'); out.close()
    print('exchanges', ctx.stats['exchanges'])
main()
