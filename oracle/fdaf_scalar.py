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


# ----------------------------------------------------------------------------------------------------------------------
# algo = 2 / 3: overlap-save partitioned-block FDAF, alternated constraint applied to the partition as it entered the block
# (DESIGN.md section 2b).  Same rules as above: plain Python, direct-sum transforms over a cosine / sine table, one bin and
# one tap at a time, nothing shared with aec_oracle.py.
# ----------------------------------------------------------------------------------------------------------------------
def _rdft(x: List[float], cos_t, sin_t) -> List[complex]:
    return [complex(sum(v * c for v, c in zip(x, cos_t[k])), -sum(v * s for v, s in zip(x, sin_t[k])))
            for k in range(len(cos_t))]


def _irdft(spec: List[complex], n: int, cos_t, sin_t) -> List[float]:
    half = n // 2
    out = []
    for i in range(n):
        s = spec[0].real + spec[half].real * (1.0 if i % 2 == 0 else -1.0)
        for k in range(1, half):
            s += 2.0 * (spec[k].real * cos_t[k][i] - spec[k].imag * sin_t[k][i])
        out.append(s / n)
    return out


def stage1_ols_scalar(far: List[float], mic: List[float], partitions: int = 4, algo: int = 2, frame: int = 512,
                      mu: float = 0.5, delta: float = None, pb_lambda: float = 0.5, a: float = 0.999, lam: float = 0.9,
                      c0: float = 1.0, eps: float = 1e-10, erle_skip_hops: int = 0) -> Tuple[List[float], List[float], float]:
    """(error signal, echo estimate, ERLE in dB) of ONE utterance.  ``algo`` 2 = NLMS step on a smoothed input power,
    3 = diagonal Kalman step.  Blocks of frame / 2 new samples; outputs cover the whole blocks only."""
    n, hop, bins = frame, frame // 2, frame // 2 + 1
    delta = 1e-6 * n if delta is None else delta
    _, cos_t, sin_t = _tables(n)
    blocks = min(len(far), len(mic)) // hop
    w = [[0j] * bins for _ in range(partitions)]            # taps, by partition
    cov = [[c0] * bins for _ in range(partitions)]
    power = [0.0] * bins                                    # smoothed input power (algo 2) / Psi (algo 3)
    spectra = []                                            # far-end spectra of the blocks so far
    err, echo = [], []
    for t in range(blocks):
        older = list(far[(t - 1) * hop:t * hop]) if t > 0 else [0.0] * hop
        newer = list(far[t * hop:(t + 1) * hop])
        spectra.append(_rdft(older + newer, cos_t, sin_t))
        past = [spectra[t - p] if t - p >= 0 else [0j] * bins for p in range(partitions)]
        yhat_spec = []
        for k in range(bins):
            acc = 0j
            for p in range(partitions):
                acc += w[p][k] * past[p][k]
            yhat_spec.append(acc)
        y = _irdft(yhat_spec, n, cos_t, sin_t)[hop:]
        d = list(mic[t * hop:(t + 1) * hop])
        e = [dv - yv for dv, yv in zip(d, y)]
        err.extend(e)
        echo.extend(y)
        e_spec = _rdft([0.0] * hop + e, cos_t, sin_t)
        # the partition whose turn it is loses the second half of its impulse response -- before this block's update
        c = t % partitions
        g = _irdft(w[c], n, cos_t, sin_t)
        w[c] = _rdft(g[:hop] + [0.0] * hop, cos_t, sin_t)
        for k in range(bins):
            if algo == 2:
                total = 0.0
                for p in range(partitions):
                    total += abs(past[p][k]) ** 2
                power[k] = pb_lambda * power[k] + (1.0 - pb_lambda) * total
                step = mu / (power[k] + delta)
                for p in range(partitions):
                    w[p][k] = w[p][k] + step * past[p][k].conjugate() * e_spec[k]
            else:
                power[k] = lam * power[k] + (1.0 - lam) * abs(e_spec[k]) ** 2
                den = power[k] + eps
                for p in range(partitions):
                    den += cov[p][k] * abs(past[p][k]) ** 2
                for p in range(partitions):
                    x2 = abs(past[p][k]) ** 2
                    gain = cov[p][k] * past[p][k].conjugate() / den
                    w[p][k] = a * (w[p][k] + gain * e_spec[k])
                    cov[p][k] = a * a * (1.0 - cov[p][k] * x2 / den) * cov[p][k] + (1.0 - a * a) * abs(w[p][k]) ** 2
    lo = erle_skip_hops * hop
    pm = sum(v * v for v in mic[lo:len(err)])
    pe = sum(v * v for v in err[lo:])
    erle = 10.0 * math.log10(max(pm, 1e-20) / max(pe, 1e-20))
    return err, echo, erle
