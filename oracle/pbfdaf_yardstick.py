"""Overlap-save constrained PBFDAF -- a YARDSTICK for the frozen STFT-domain recurrence, nothing more.
TEST / ANALYSIS INFRASTRUCTURE ONLY (never imported by the product; not on any parity path).

The classical partitioned-block frequency-domain adaptive filter (rectangular blocks of H new samples, FFT length 2H,
P partitions, linear convolution by overlap-save, gradient constraint through an inverse / forward transform pair):

    X_p(t)  = FFT_2H[ x block t-p-1 , x block t-p ]                      p = 0 .. P-1
    yhat(t) = last H samples of IFFT( sum_p W_p X_p(t) )
    e(t)    = d(t) - yhat(t)                E(t) = FFT_2H[ 0_H , e(t) ]
    Pw      = lam Pw + (1 - lam) sum_p |X_p(t)|^2
    W_p    += mu * FFT_2H[ first H samples of IFFT( conj(X_p) E / (Pw + delta) ) , 0_H ]

It models the same P*H-sample echo tail as the frozen recurrence with P partitions of hop H, but
as an exact linear convolution: no Hann analysis window, no cross-band leakage.  tools/erle_yardstick.py reports the
ERLE of both on the SURVEY 8d single-talk / double-talk sets; DESIGN.md section 2 quotes the table.
"""
from __future__ import annotations

import numpy as np


def pbfdaf(far: np.ndarray, mic: np.ndarray, partitions: int = 4, hop: int = 256, mu: float = 0.5,
           delta: float = 1e-6, lam: float = 0.9, constrained: bool = True):
    """One utterance.  Returns (err, yhat), each (len // hop) * hop samples, time-aligned with the inputs."""
    far = np.asarray(far, dtype=np.float64)
    mic = np.asarray(mic, dtype=np.float64)
    H, P = hop, partitions
    nblk = min(len(far), len(mic)) // H
    K = H + 1
    W = np.zeros((P, K), dtype=np.complex128)
    Xh = np.zeros((P, K), dtype=np.complex128)        # X_p, p = 0 newest
    pw = np.zeros(K)
    prev = np.zeros(H)
    err = np.zeros(nblk * H)
    yh = np.zeros(nblk * H)
    for t in range(nblk):
        cur = far[t * H:(t + 1) * H]
        Xh = np.roll(Xh, 1, axis=0)
        Xh[0] = np.fft.rfft(np.concatenate([prev, cur]))
        prev = cur
        y = np.fft.irfft((W * Xh).sum(axis=0), n=2 * H)[H:]
        e = mic[t * H:(t + 1) * H] - y
        err[t * H:(t + 1) * H] = e
        yh[t * H:(t + 1) * H] = y
        E = np.fft.rfft(np.concatenate([np.zeros(H), e]))
        pw = lam * pw + (1.0 - lam) * (np.abs(Xh) ** 2).sum(axis=0)
        G = np.conj(Xh) * (E / (pw + delta))
        if constrained:
            g = np.fft.irfft(G, n=2 * H, axis=1)
            g[:, H:] = 0.0
            G = np.fft.rfft(g, axis=1)
        W += mu * G
    return err, yh
