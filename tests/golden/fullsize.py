"""Seeded inputs and pinned positions of the full-size (10 s, 16 kHz) reference fixture.

Shared by tests/golden/make_golden.py (which imports the reference and writes reference_full_size.npz) and by the
tests (which must not need the reference): only summaries of the reference's outputs are committed, the inputs are
regenerated from the seed."""
import numpy as np


def full_size_inputs(seed=424242, B=2, L=160000):
    rng = np.random.default_rng(seed)
    mic = (0.2 * rng.standard_normal((B, L)) + 0.01).astype(np.float32)
    ref = (0.2 * rng.standard_normal((B, L)) - 0.02).astype(np.float32)
    return mic, ref


def full_size_positions(shape, n=256, seed=7):
    """Fixed pseudo-random flat positions at which single values of an output are pinned."""
    return np.random.default_rng(seed).integers(0, int(np.prod(shape)), size=n)
