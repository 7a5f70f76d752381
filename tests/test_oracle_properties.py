"""CPU: oracle-free properties of the builder-authored recurrence and the C restatement against the
numpy oracle.  (FDAF parity is UNPINNED by the reference: it has no stage-1 filter.)"""
import numpy as np
import pytest

from acoustic_echo_cancellation_b200 import synth
from oracle import aec_oracle as O
from oracle import c_oracle as CO


def test_zero_far_end_passes_microphone_through():
    rng = np.random.default_rng(1)
    mic = (0.2 * rng.standard_normal((2, 8192))).astype(np.float32)
    for algo in (O.ALGO_NLMS, O.ALGO_KALMAN):
        r = O.stage1(np.zeros_like(mic), mic, O.AecConfig(algo=algo))
        assert np.abs(r["err"] - mic[:, :r["err"].shape[1]]).max() < 1e-6
        assert np.abs(r["echo"]).max() == 0


def test_echo_plus_error_is_stft_roundtrip_of_mic():
    d = synth.make_batch(3, 2, 16000)
    r = O.stage1(d["far"], d["mic"], O.AecConfig())
    n = r["err"].shape[1]
    assert np.abs(r["echo"] + r["err"] - d["mic"][:, :n]).max() < 1e-6


@pytest.mark.parametrize("algo,floor", [(O.ALGO_NLMS, 10.0), (O.ALGO_KALMAN, 10.0)])
def test_single_talk_erle_floor(algo, floor):
    d = synth.make_batch(0, 2, 160000)
    r = O.stage1(d["far"], d["mic"], O.AecConfig(algo=algo), erle_skip=32000)
    assert (r["erle_db"] > floor).all(), r["erle_db"]


def test_nlms_is_linear_in_the_microphone():
    d = synth.make_batch(5, 1, 8192)
    rng = np.random.default_rng(2)
    m2 = (0.1 * rng.standard_normal((1, 8192))).astype(np.float32)
    cfg = O.AecConfig()
    a = O.stage1(d["far"], d["mic"], cfg)["err"]
    b = O.stage1(d["far"], m2, cfg)["err"]
    c = O.stage1(d["far"], 0.5 * d["mic"].astype(np.float64) - 2.0 * m2, cfg)["err"]
    assert np.abs(c - (0.5 * a - 2.0 * b)).max() < 1e-9


def test_float32_drift_is_far_below_tolerance():
    d = synth.make_batch(0, 1, 160000)
    for algo in (O.ALGO_NLMS, O.ALGO_KALMAN):
        cfg = O.AecConfig(algo=algo)
        a = O.stage1(d["far"], d["mic"], cfg)["err"]
        b = O.stage1(d["far"], d["mic"], cfg, dtype=np.float32)["err"]
        assert np.abs(a - b).max() < 2e-5


@pytest.mark.parametrize("algo,P,L", [(0, 4, 16123), (1, 4, 16000), (1, 16, 12000), (0, 8, 8000), (0, 1, 4000)])
def test_c_oracle_matches_numpy_oracle(algo, P, L):
    cfg = O.AecConfig(partitions=P, algo=algo)
    d = synth.make_batch(0, 3, L, rir_len=P * 256)
    ns = np.array([L, L - 1, L - 777])
    r = O.stage1(d["far"], d["mic"], cfg, n_samples=ns, erle_skip=8 * 256)
    c = CO.stage1(d["far"], d["mic"], cfg, n_samples=ns, erle_skip_hops=8)
    n = r["err"].shape[1]
    assert np.abs(c["err"][:, :n] - r["err"]).max() < 1e-5
    assert np.abs(c["echo"][:, :n] - r["echo"]).max() < 1e-5
    assert np.abs(c["erle_db"] - r["erle_db"]).max() < 1e-3
    assert (c["err"][:, n:] == 0).all()
    for b in range(3):
        m = (O.n_frames(int(ns[b])) - 1) * 256
        assert (c["err"][b, m:] == 0).all()


@pytest.mark.parametrize("algo,P,L,frame", [(2, 4, 16123, 512), (3, 4, 16000, 512), (3, 16, 12000, 512), (2, 8, 8000, 512),
                                            (2, 1, 4000, 512), (3, 8, 20000, 1024), (2, 2, 300, 512)])
def test_c_oracle_overlap_save_matches_numpy_oracle(algo, P, L, frame):
    """the port runs the overlap-save filters eight utterances abreast (SIMD lanes = utterances): 11 ragged utterances
    = one full group and a partial one, every lane against the float64 oracle run on its utterance alone"""
    H = frame // 2
    cfg = O.AecConfig(frame=frame, partitions=P, algo=algo, delta=1e-6 * frame)
    B = 11
    d = synth.make_batch(0, B, L, rir_len=P * H)
    ns = np.array([L, L - 1, max(L - 777, 0)] + [max(L - 100 * i, 0) for i in range(B - 3)])
    r = O.stage1(d["far"], d["mic"], cfg, n_samples=ns, erle_skip=8 * H)
    c = CO.stage1(d["far"], d["mic"], cfg, n_samples=ns, erle_skip_hops=8)
    n = r["err"].shape[1]
    assert np.abs(c["err"][:, :n] - r["err"]).max() < 2e-5
    assert np.abs(c["echo"][:, :n] - r["echo"]).max() < 2e-5
    assert np.abs(c["erle_db"] - r["erle_db"]).max() < 1e-3
    for b in range(B):
        m = int(ns[b]) // H * H
        assert (c["err"][b, m:] == 0).all() and (c["echo"][b, m:] == 0).all()
    # a lane does not see its neighbours: utterance 4 alone gives the same bits
    alone = CO.stage1(d["far"][4:5], d["mic"][4:5], cfg, n_samples=ns[4:5], erle_skip_hops=8)
    assert np.array_equal(alone["err"][0], c["err"][4]) and alone["erle_db"][0] == c["erle_db"][4]


def test_ragged_equals_alone():
    d = synth.make_batch(7, 2, 6000)
    ns = np.array([6000, 3333])
    r = O.stage1(d["far"], d["mic"], O.AecConfig(), n_samples=ns)
    alone = O.stage1(d["far"][1:2, :3333], d["mic"][1:2, :3333], O.AecConfig())
    m = alone["err"].shape[1]
    assert np.array_equal(r["err"][1, :m], alone["err"][0])
    assert (r["err"][1, m:] == 0).all()


def test_flop_model_matches_survey_table():
    assert O.flops_per_frame(O.AecConfig(partitions=4, algo=0)) == pytest.approx(56652)
    assert O.flops_per_frame(O.AecConfig(partitions=16, algo=1)) == pytest.approx(167419)
    assert O.flops_per_frame(O.AecConfig(frame=1024, partitions=8, algo=0)) == pytest.approx(153740)
    # overlap-save filters (builder's count, same conventions): 12.5 N log2 N + 2 N + the recurrence terms
    assert O.flops_per_frame(O.AecConfig(partitions=4, algo=O.ALGO_PBFDAF)) == pytest.approx(78156)
    assert O.flops_per_frame(O.AecConfig(partitions=4, algo=O.ALGO_PBFKF)) == pytest.approx(93319)
    import bench
    for algo in (0, 1, 2, 3):
        assert bench.flops_per_frame(512, 4, algo) == pytest.approx(O.flops_per_frame(O.AecConfig(partitions=4, algo=algo)))


def test_overlap_save_pbfdaf_oracle_properties():
    """algo = 2 (oracle only; the GPU parity tests compare the kernel with it): linear in the microphone, identity for a
    silent far end, float32 tracks float64, and it reaches the -40 dB noise floor where the STFT-domain NLMS stays near
    13 dB (the reason it exists)"""
    from acoustic_echo_cancellation_b200 import synth

    d = synth.make_utterance(3, 48000, rir_len=1024)
    cfg = O.AecConfig(partitions=4, algo=O.ALGO_PBFDAF)
    far, mic = d["far"].astype(np.float64), d["mic"].astype(np.float64)
    e, yh = O.pbfdaf_ols(far, mic, cfg)
    assert e.shape == yh.shape == (48000 // 256 * 256,)
    assert np.abs(e + yh - mic[:e.size]).max() < 1e-12                       # e = d - y by construction
    e0, y0 = O.pbfdaf_ols(np.zeros_like(far), mic, cfg)
    assert np.array_equal(e0, mic[:e0.size]) and not y0.any()                # silent far end: identity
    rng = np.random.default_rng(0)
    mic2 = 0.05 * rng.standard_normal(mic.size)
    ea, _ = O.pbfdaf_ols(far, mic2, cfg)
    ec, _ = O.pbfdaf_ols(far, 0.5 * mic - 2.0 * mic2, cfg)
    assert np.abs(ec - (0.5 * e - 2.0 * ea)).max() < 1e-10                   # the step does not depend on the microphone
    e32, _ = O.pbfdaf_ols(far, mic, cfg, dtype=np.float32)
    assert np.abs(e32 - e).max() < 1e-5
    # algo 3 (Kalman step): identity for a silent far end, float32 tracks float64, e + y = d
    kcfg = O.AecConfig(partitions=4, algo=O.ALGO_PBFKF)
    ek, yk = O.pbfdaf_ols(far, mic, kcfg)
    assert np.abs(ek + yk - mic[:ek.size]).max() < 1e-12
    ek0, yk0 = O.pbfdaf_ols(np.zeros_like(far), mic, kcfg)
    assert np.array_equal(ek0, mic[:ek0.size]) and not yk0.any()
    ek32, _ = O.pbfdaf_ols(far, mic, kcfg, dtype=np.float32)
    assert np.abs(ek32 - ek).max() < 1e-5
    lo = 24000
    assert 10 * np.log10((mic[lo:ek.size] ** 2).sum() / (ek[lo:] ** 2).sum()) > 20.0      # 25 dB after 1.5 s (35 after 4)
    erle = 10 * np.log10((mic[lo:e.size] ** 2).sum() / (e[lo:] ** 2).sum())
    r = O.stage1(d["far"][None], d["mic"][None], O.AecConfig(partitions=4, algo=O.ALGO_NLMS))
    erle_stft = 10 * np.log10((mic[lo:e.size] ** 2).sum() / (r["err"][0][lo:e.size] ** 2).sum())
    assert erle > 33.0 and erle_stft < 20.0
