"""qgate_b200 — B200-native state-vector engine behind qgate's simulator.cuda() API.

    import qgate_b200 as qgate
    from qgate_b200.script import *
    sim = qgate.simulator.cuda(dtype=np.float64)
"""
from . import model
from .model import prefs
from . import script
from . import simulator
from .dump import dump

__version__ = '0.1.0'
