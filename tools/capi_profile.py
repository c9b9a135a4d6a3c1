#!/usr/bin/env python
"""Run a script (bench.py, run_configs.py ...) with the time spent in every C-ABI function reached through
CApi.call accumulated, and printed at exit on rank 0:  python tools/capi_profile.py bench.py --gpus 2 ..."""
import atexit
import os
import runpy
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from qgate_b200 import _capi  # noqa: E402

acc, cnt = {}, {}
orig = _capi.CApi.call


def call(self, name, *args):
    t0 = time.perf_counter()
    try:
        return orig(self, name, *args)
    finally:
        acc[name] = acc.get(name, 0.) + time.perf_counter() - t0
        cnt[name] = cnt.get(name, 0) + 1


_capi.CApi.call = call


def report():
    if os.environ.get('RANK', '0') != '0':
        return
    for name, t in sorted(acc.items(), key=lambda kv: -kv[1])[:12]:
        sys.stderr.write('[capi] %-40s %8.1f ms in %6d calls\n' % (name, 1e3 * t, cnt[name]))


atexit.register(report)
script = sys.argv[1]
sys.argv = sys.argv[1:]
runpy.run_path(os.path.join(REPO, script) if not os.path.isabs(script) else script, run_name='__main__')
