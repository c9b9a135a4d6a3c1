"""Build qgate_b200/lib/libqgate_b200.so for sm_100a with nvcc (in-tree, no JIT cache).

    python qgate_b200/csrc/build.py [--force]

Objects are rebuilt only when a source or header is newer; the .so is git-ignored but
travels to the GPU box with the repo snapshot.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_DIR = os.path.join(os.path.dirname(HERE), 'lib')
OBJ_DIR = os.path.join(HERE, 'build')
LIB = os.path.join(LIB_DIR, 'libqgate_b200.so')

# (source, object name, extra flags): the fused-pass kernel file is compiled twice, its complex128 and
# complex64 instantiations in parallel (4 min -> 2.5 min of wall clock)
SOURCES = [('engine.cu', 'engine.cu', []), ('dist.cu', 'dist.cu', []),
           ('kernels_tma.cu', 'kernels_tma_f64.cu', ['-DQGB_TMA_PART=0']),
           ('kernels_tma.cu', 'kernels_tma_f32.cu', ['-DQGB_TMA_PART=1']),
           ('kernels_ops.cu', 'kernels_ops.cu', []), ('planner.cpp', 'planner.cpp', []),
           ('gate_matrix.cpp', 'gate_matrix.cpp', [])]
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
         '-Xcompiler', '-fPIC,-Wall,-Wno-unused-function', '-Xptxas', '-v']
if os.environ.get('QGB_PHASE_TIMING'):      # diagnostic build: per-phase cycle counters in the pass kernel
    FLAGS.append('-DQGB_PHASE_TIMING')


def _newest_header():
    times = [os.path.getmtime(os.path.join(HERE, f)) for f in os.listdir(HERE) if f.endswith('.h')]
    times.append(os.path.getmtime(os.path.join(HERE, '..', '..', 'include', 'qgate_b200.h')))
    return max(times)


def _compile(entry, force, log):
    src, name, extra = entry
    obj = os.path.join(OBJ_DIR, name + '.o')
    src_path = os.path.join(HERE, src)
    if not force and os.path.exists(obj) and \
            os.path.getmtime(obj) >= max(os.path.getmtime(src_path), _newest_header()):
        return obj, False
    cmd = [NVCC] + FLAGS + extra + ['-c', src_path, '-o', obj]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    with open(os.path.join(OBJ_DIR, name + '.log'), 'w') as f:
        f.write(res.stdout)
    if res.returncode != 0:
        raise RuntimeError('nvcc failed for {}:\n{}'.format(src, res.stdout))
    return obj, True


def build(force=False, verbose=False):
    os.makedirs(LIB_DIR, exist_ok=True)
    os.makedirs(OBJ_DIR, exist_ok=True)
    sources = [s for s in SOURCES if os.path.exists(os.path.join(HERE, s[0]))]
    with ThreadPoolExecutor(max_workers=len(sources)) as ex:
        results = list(ex.map(lambda s: _compile(s, force, verbose), sources))
    objs = [o for o, _ in results]
    if any(changed for _, changed in results) or not os.path.exists(LIB):
        cmd = [NVCC, '-shared', '-gencode', 'arch=compute_100a,code=sm_100a'] + objs + \
              ['-o', LIB, '-Xlinker', '--no-undefined', '-ldl']
        res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if res.returncode != 0:
            raise RuntimeError('link failed:\n' + res.stdout)
    if verbose:
        print('built', LIB)
    return LIB


if __name__ == '__main__':
    build(force='--force' in sys.argv, verbose=True)
