"""Parity hardening for the (unpinned) FDAF recurrence: a second restatement of DESIGN.md section 2, written one bin
and one tap at a time in plain Python with direct-sum transforms and no helper shared with the numpy oracle
(oracle/fdaf_scalar.py), must agree with the numpy oracle (CPU) and with the CUDA path (GPU) -- 3 seeds x
{NLMS, Kalman} for the STFT-domain recurrence, 2 seeds x {NLMS, Kalman step} for the overlap-save filter.  A mistake the oracle and the kernels share because one hand wrote both would show up here."""
import numpy as np
import pytest

from acoustic_echo_cancellation_b200 import synth
from oracle import aec_oracle as O
from oracle import fdaf_scalar as S

L = 2500          # not a multiple of the hop; 10 frames
CASES = [(seed, algo) for seed in (0, 1, 2) for algo in (0, 1)] + [(seed, algo) for seed in (0, 2) for algo in (2, 3)]
_cache = {}


def _scalar(seed, algo, P=4):
    key = (seed, algo, P)
    if key not in _cache:
        d = synth.make_utterance(100 + seed, L, rir_len=512, double_talk=(seed == 2))
        far, mic = d["far"].astype(np.float64), d["mic"].astype(np.float64)
        run = S.stage1_scalar if algo < 2 else S.stage1_ols_scalar      # algos 2 / 3: overlap-save filter
        err, echo, erle = run(list(far), list(mic), partitions=P, algo=algo, erle_skip_hops=2)
        _cache[key] = (d, np.array(err), np.array(echo), erle)
    return _cache[key]


@pytest.mark.parametrize("seed,algo", CASES)
def test_scalar_restatement_agrees_with_numpy_oracle(seed, algo):
    d, err, echo, erle = _scalar(seed, algo)
    ref = O.stage1(d["far"], d["mic"], O.AecConfig(partitions=4, algo=algo), erle_skip=2 * 256)
    assert err.shape == ref["err"][0].shape == (9 * 256,)
    assert np.abs(err - ref["err"][0]).max() <= 1e-10
    assert np.abs(echo - ref["echo"][0]).max() <= 1e-10
    assert abs(erle - ref["erle_db"][0]) <= 1e-8
    # the filter does something on this input (not a vacuous comparison)
    assert np.abs(echo).max() > 1e-3 and np.abs(err - d["mic"][:err.size]).max() > 1e-3


@pytest.mark.gpu
@pytest.mark.parametrize("seed,algo", CASES)
def test_scalar_restatement_agrees_with_cuda_path(seed, algo):
    import torch

    import acoustic_echo_cancellation_b200 as A

    d, err, echo, erle = _scalar(seed, algo)
    cfg = A.Stage1Config(partitions=4, algo=algo, erle_skip_hops=2)
    far = torch.from_numpy(d["far"][None]).cuda()
    mic = torch.from_numpy(d["mic"][None]).cuda()
    e, yh, g_erle = A.stage1_aec(far, mic, cfg, return_echo=True, return_erle=True)
    torch.cuda.synchronize()
    e, yh = e.cpu().numpy()[0], yh.cpu().numpy()[0]
    assert np.abs(e[:err.size] - err).max() <= 1e-4 and (e[err.size:] == 0).all()
    assert np.abs(yh[:echo.size] - echo).max() <= 1e-4
    assert abs(float(g_erle[0]) - erle) <= 0.05
