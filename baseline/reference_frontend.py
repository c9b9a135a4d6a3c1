"""Loader of the UNMODIFIED reference package (shinmorino/qgate 0.2.2) for the two places that run
the reference's own Python front end on top of this repository's runtime:

  tests/test_reference_frontend.py   the reference's own unittest suite, `...CUDA` classes
  bench.py (e2e leg)                 qgate.simulator.cuda().run(circuit) after
                                     qgate_b200.install.install(qgate)

Where it comes from: a scratch build of /root/reference (build container), or baseline/_ref/ — the
copy scripts/install_reference.sh leaves there (git-ignored; it travels to the GPU box with the
gpurun snapshot, where /root/reference does not exist).  Nothing of it is tracked in this repository.
The OpenQASM importer is stubbed: it needs PLY, which the image does not have, and is out of scope
(SURVEY.md section 2.1 row 10).
"""
import os
import subprocess
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
INSTALLED = os.path.join(HERE, '_ref')
REFERENCE = '/root/reference'
SCRATCH = '/tmp/qgate_ref'


def root():
    """Directory holding `qgate/` (+ `tests/`, `examples/`) of the reference, or None."""
    if os.path.exists(os.path.join(INSTALLED, 'qgate', 'simulator', 'cpuext.so')):
        return INSTALLED
    if os.path.isdir(os.path.join(REFERENCE, 'qgate')):
        if not os.path.exists(os.path.join(SCRATCH, 'qgate', 'simulator', 'cpuext.so')):
            subprocess.check_call('rm -rf {0} && cp -r {1} {0} && chmod -R u+w {0}'.format(SCRATCH, REFERENCE),
                                  shell=True)
            subprocess.check_call('python3 incpathgen.py > incpath && make ../glue.so ../cpuext.so -j8',
                                  shell=True, cwd=os.path.join(SCRATCH, 'qgate', 'simulator', 'src'),
                                  stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        return SCRATCH
    return None


def available():
    return root() is not None


def load():
    """import qgate (the reference) and return the package."""
    where = root()
    if where is None:
        raise ImportError('the reference package is neither under baseline/_ref nor at /root/reference')
    if where not in sys.path:
        sys.path.insert(0, where)
    sys.modules.setdefault('qgate.openqasm', types.ModuleType('qgate.openqasm'))
    import numpy as np
    if not hasattr(np, 'int'):
        np.int = int          # the reference (and its tests) use the alias NumPy 1.24 removed
    import qgate
    return qgate
