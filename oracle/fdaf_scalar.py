"""Second, independently structured restatement of DESIGN.md section 2.  TEST INFRASTRUCTURE ONLY.

Written from the recurrence as DESIGN.md states it, one bin and one tap at a time in plain Python complex / float
arithmetic: no numpy broadcasting over bins, no helper shared with ``aec_oracle.py`` (not even its STFT -- the
transforms here are direct sums over a cosine / sine table, the way the reference's conv kernels are defined,
Stage2_lhm/scripts/network/attention_ccrn.py:8-25, and the overlap-add follows :82-101 literally: add every
windowed inverse frame into a long buffer, add the squared windows into a second one, divide, trim).  Its only
purpose is to catch a mistake that the numpy oracle and the kernels might share because one hand wrote both:
``tests/test_scalar_restatement.py`` checks it against the numpy oracle on the CPU and against the GPU.
Slow by design (seconds per second of audio); use it on short signals.

PARITY UNPINNED, like everything that concerns the FDAF recurrence: the reference has no stage-1 filter.
"""
from __future__ import annotations

import math
from typing import List, Tuple


def _tables(n: int):
    win = [0.5 - 0.5 * math.cos(2.0 * math.pi * i / n) for i in range(n)]        # periodic Hann, attention_ccrn.py:12
    cos_t = [[math.cos(2.0 * math.pi * k * i / n) for i in range(n)] for k in range(n // 2 + 1)]
    sin_t = [[math.sin(2.0 * math.pi * k * i / n) for i in range(n)] for k in range(n // 2 + 1)]
    return win, cos_t, sin_t


def _analysis(x: List[float], n: int, tabs) -> List[List[complex]]:
    """frames of the zero-padded signal times window, real part sum x w cos, imaginary part -sum x w sin"""
    win, cos_t, sin_t = tabs
    hop = n // 2
    pad = n - hop
    xp = [0.0] * pad + list(x) + [0.0] * pad
    t_count = (len(xp) - n) // hop + 1 if len(xp) >= n else 0
    out = []
    for t in range(t_count):
        seg = [xp[t * hop + i] * win[i] for i in range(n)]
        row = []
        for k in range(n // 2 + 1):
            re = sum(s * c for s, c in zip(seg, cos_t[k]))
            im = -sum(s * c for s, c in zip(seg, sin_t[k]))
            row.append(complex(re, im))
        out.append(row)
    return out


def _synthesis(spec: List[List[complex]], n: int, tabs) -> List[float]:
    win, cos_t, sin_t = tabs
    hop = n // 2
    t_count = len(spec)
    if t_count == 0:
        return []
    total = (t_count - 1) * hop + n
    acc = [0.0] * total
    norm = [0.0] * total
    half = n // 2
    for t, row in enumerate(spec):
        for i in range(n):
            # inverse real DFT: weights 1/n at DC and Nyquist, 2/n elsewhere
            s = row[0].real + row[half].real * (1.0 if i % 2 == 0 else -1.0)
            for k in range(1, half):
                s += 2.0 * (row[k].real * cos_t[k][i] - row[k].imag * sin_t[k][i])
            acc[t * hop + i] += (s / n) * win[i]
            norm[t * hop + i] += win[i] * win[i]
    y = [a / (c + 1e-8) for a, c in zip(acc, norm)]
    return y[n - hop: total - (n - hop)]


def stage1_scalar(far: List[float], mic: List[float], partitions: int = 4, algo: int = 0, frame: int = 512,
                  mu: float = 0.5, delta: float = None, a: float = 0.999, lam: float = 0.9, c0: float = 1.0,
                  eps: float = 1e-10, erle_skip_hops: int = 0) -> Tuple[List[float], List[float], float]:
    """(error signal, echo estimate, ERLE in dB) of ONE utterance.  ``algo`` 0 = NLMS, 1 = Kalman."""
    n, hop, bins = frame, frame // 2, frame // 2 + 1
    delta = 1e-6 * n if delta is None else delta
    tabs = _tables(n)
    xs = _analysis(far, n, tabs)
    ys = _analysis(mic, n, tabs)
    frames = len(xs)
    e_spec = [[0j] * bins for _ in range(frames)]
    h_spec = [[0j] * bins for _ in range(frames)]
    for k in range(bins):                       # bins are independent: run each one through time on its own
        w = [0j] * partitions
        cov = [c0] * partitions
        psi = 0.0
        for t in range(frames):
            past = [xs[t - p][k] if t - p >= 0 else 0j for p in range(partitions)]
            yhat = 0j
            for p in range(partitions):
                yhat += w[p] * past[p]
            e = ys[t][k] - yhat
            if algo == 0:
                power = 0.0
                for p in range(partitions):
                    power += abs(past[p]) ** 2
                g = mu / (power + delta)
                for p in range(partitions):
                    w[p] = w[p] + g * past[p].conjugate() * e
            else:
                psi = lam * psi + (1.0 - lam) * abs(e) ** 2
                d = psi + eps
                for p in range(partitions):
                    d += cov[p] * abs(past[p]) ** 2
                for p in range(partitions):
                    x2 = abs(past[p]) ** 2
                    gain = cov[p] * past[p].conjugate() / d
                    w[p] = a * (w[p] + gain * e)
                    cov[p] = a * a * (1.0 - cov[p] * x2 / d) * cov[p] + (1.0 - a * a) * abs(w[p]) ** 2
            e_spec[t][k] = e
            h_spec[t][k] = yhat
    err = _synthesis(e_spec, n, tabs)
    echo = _synthesis(h_spec, n, tabs)
    lo = erle_skip_hops * hop
    pm = sum(v * v for v in mic[lo:len(err)])
    pe = sum(v * v for v in err[lo:])
    erle = 10.0 * math.log10(max(pm, 1e-20) / max(pe, 1e-20))
    return err, echo, erle
