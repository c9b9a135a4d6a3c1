#!/usr/bin/env python
"""The drop-in route: the reference package (shinmorino/qgate 0.2.x) keeps its whole Python front end,
`qgate.simulator.cuda()` is served by libqgate_b200.so.  `import qgate` must work (a checkout of the
reference on PYTHONPATH, or baseline/_ref of this repository)."""
import math
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
try:
    import qgate
except ImportError:
    from baseline import reference_frontend
    qgate = reference_frontend.load()
import qgate_b200.install

qgate_b200.install.install(qgate)              # sets qgate.simulator.cudaruntime (and qgate.openqasm if PLY is missing)
from qgate.script import H, ctrl, X, U1, new_qregs, new_references, measure   # noqa: E402

n = 24
q = new_qregs(n)
refs = new_references(n)
circuit = [H(q[0])] + [ctrl(q[i]).X(q[i + 1]) for i in range(n - 1)]          # GHZ
circuit += [ctrl(q[0]).U1(math.pi / 3)(q[n - 1])]
circuit += [measure(r, x) for r, x in zip(refs, q)]
sim = qgate.simulator.cuda(dtype=np.float64)
sim.run(circuit)
print('observed:', format(sim.obs(refs).int, '0{}b'.format(n)))              # all zeros or all ones
sim.terminate()
