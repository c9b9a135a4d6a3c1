"""CPU suite: the oracle-free checks of run_configs.py (BASELINE configs 3-5) hold on the reference
CPU runtime at small sizes — closed-form Grover probability, phase-estimation peak and closed-form
QFT amplitudes — so a failure of `run_configs.py` on the GPU box points at the engine, not at the
check.  The GPU entry point itself (`simulator.cuda`) is replaced by the oracle runtime here."""
import types

import numpy as np
import pytest

import qgate_b200
import run_configs


class _NoTorch:
    class cuda:
        @staticmethod
        def synchronize():
            pass


@pytest.fixture()
def oracle_as_cuda(ref_runtime, monkeypatch):
    monkeypatch.setattr(qgate_b200.simulator, 'cuda',
                        lambda **prefs: qgate_b200.simulator.with_runtime(ref_runtime.module, **prefs))


def _args(**kw):
    return types.SimpleNamespace(iterations=8, shots=100000, dtype='f64', **kw)


def test_grover_check(oracle_as_cuda):
    out = run_configs.run_grover(_args(qubits=12), _NoTorch, 1, 0)
    assert out['ok'], out
    assert abs(out['p_marked'] - out['p_marked_closed_form']) < 1e-12


def test_phase_estimation_check(oracle_as_cuda):
    out = run_configs.run_pe(_args(qubits=13), _NoTorch, 1, 0)
    assert out['ok'], out
    assert out['peak'] == out['peak_expected'] and abs(out['estimate'] - 0.1) < 2. ** -12


@pytest.mark.parametrize('n', (9, 14))
def test_qft_check(oracle_as_cuda, n):
    out = run_configs.run_qft(_args(qubits=n), _NoTorch, 1, 0)
    assert out['ok'], out
    assert out['amplitude_rel_err_vs_closed_form'] < 1e-12
