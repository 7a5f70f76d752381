"""CPU: the oracle's reference-pinned parts against golden vectors produced by importing the
reference's own operators (tests/golden/make_golden.py).  Tolerances: the reference computes a
float32 dense convolution, the oracle a float64 FFT -> ~1e-5 abs on spectra of 0.3-sigma noise."""
import numpy as np
import pytest

from oracle import aec_oracle as O


@pytest.mark.parametrize("i", range(5))
def test_stft_matches_reference_convstft(golden, i):
    x, s = golden[f"x512_{i}"], golden[f"stft512_{i}"]
    so = O.stft(x)
    assert so.shape == s.shape
    assert np.abs(so - s).max() < 5e-5


@pytest.mark.parametrize("i", range(5))
def test_istft_matches_reference_convistft(golden, i):
    s, y = golden[f"stft512_{i}"], golden[f"istft512_{i}"]
    yo = O.istft(s.astype(np.float64))
    assert yo.shape == y.shape
    if y.size:
        assert np.abs(yo - y).max() < 5e-6


def test_istft_free_spectrum_ignores_dc_nyquist_imag(golden):
    yo = O.istft(golden["spec_free"].astype(np.float64))
    assert np.abs(yo - golden["istft_free"]).max() < 5e-6


def test_frame_1024(golden):
    so = O.stft(golden["x1024"], 1024, 512)
    assert so.shape == golden["stft1024"].shape
    assert np.abs(so - golden["stft1024"]).max() < 1e-4
    yo = O.istft(golden["stft1024"].astype(np.float64), 1024, 512)
    assert np.abs(yo - golden["istft1024"]).max() < 5e-6


def test_frame_count_and_countframes_quirk(golden):
    for n, t, cf in zip(golden["frame_count_L"], golden["frame_count_T"], golden["countFrames_ref"]):
        assert O.n_frames(int(n)) == int(t)
        assert O.count_frames_reference(int(n), 512, 256) == int(cf)
    # the reference helper is one short of the STFT for a 10 s signal (tools.py:30-32)
    assert O.n_frames(160000) == 626 and O.count_frames_reference(160000, 512, 256) == 625


def test_erb_filterbank_bit_exact(golden):
    assert np.array_equal(O.erb_filterbank(), golden["erb"])


def test_product_erb_filterbank_bit_exact(golden):
    """the PRODUCT's own bank (spectral.erb_filterbank, what a caller passes to aec_features), not only the
    oracle's copy, against the array the reference's EquivalentRectangularBandwidth(...).filters holds"""
    from acoustic_echo_cancellation_b200 import spectral

    assert np.array_equal(spectral.erb_filterbank(), golden["erb"])
    assert spectral.erb_filterbank().dtype == golden["erb"].dtype


def test_feature_front_end(golden):
    f = O.stage2_features(golden["feat_mic"], golden["feat_ref"], golden["erb"])
    ref = golden["feat"]
    assert f.shape == ref.shape
    assert np.abs(f - ref).max() < 2e-4 * max(1.0, np.abs(ref).max())


def test_stage2_little_net_inference_matches_reference_module():
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_stage2.npz"))
    w = {k[2:]: g[k] for k in g.files if k.startswith("w_")}
    out = O.stage2_little_net(g["mic"], g["ref"], g["erb"], w)
    assert out.shape == g["out_wav"].shape
    assert np.abs(out - g["out_wav"]).max() < 2e-4 * max(1.0, np.abs(g["out_wav"]).max())


# ---- BASELINE.json's utterance length (10 s) against summaries of the reference's own outputs -------------------
def _full():
    import os
    import sys
    here = os.path.join(os.path.dirname(__file__), "golden")
    sys.path.insert(0, here)
    from fullsize import full_size_inputs, full_size_positions
    return np.load(os.path.join(here, "reference_full_size.npz")), full_size_inputs, full_size_positions


def _check_summary(name, arr, g, pos_fn, tol_val, sums=()):
    assert list(arr.shape) == list(g[name + "_shape"])
    ref = g[name + "_vals"]
    got = arr.ravel()[pos_fn(arr.shape)]
    assert np.abs(got - ref).max() <= tol_val * max(1.0, np.abs(ref).max())
    for key, fn in sums:
        r = g[key]
        assert np.abs(fn(arr.astype(np.float64)) - r).max() <= 1e-4 * max(1.0, np.abs(r).max())


def test_full_size_stft_istft_features_against_reference_summaries():
    g, inputs, pos = _full()
    mic, ref = inputs()
    s = O.stft(mic)                                                        # [2, 514, 626]
    _check_summary("stft", s, g, pos, 2e-5, [("stft_sum_t", lambda a: a.sum(axis=2)), ("stft_pow_c", lambda a: (a ** 2).sum(axis=1))])
    y = O.istft(s)
    _check_summary("istft", y.reshape(g["istft_shape"]), g, pos, 2e-5, [("istft_pow", lambda a: (a ** 2).sum(axis=(1, 2)))])
    f = O.stage2_features(mic, ref, O.erb_filterbank())
    _check_summary("feat", f, g, pos, 2e-4, [("feat_sum_t", lambda a: a.sum(axis=1))])


def test_full_size_little_net_against_reference_summary():
    import os
    g, inputs, pos = _full()
    mic, ref = inputs()
    w = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_stage2.npz"))
    out = O.stage2_little_net(mic, ref, O.erb_filterbank(), {k[2:]: w[k] for k in w.files if k.startswith("w_")})
    _check_summary("net", out, g, pos, 3e-4, [("net_pow", lambda a: (a ** 2).sum(axis=-1).reshape(-1))])
