"""GPU (-m gpu): parity of the CUDA path, called through the C ABI, against the oracle.

Bar (north star): max-abs error <= 1e-4 on the time-domain error signal, ERLE within 0.05 dB.
STFT / iSTFT / feature front end are additionally checked against golden vectors produced by the
reference's own operators.  The FDAF recurrence has NO reference implementation (parity unpinned):
those cases compare against the builder-authored float64 oracle and against oracle-free properties
at BASELINE.json's full sizes.
"""
import ctypes as C

import numpy as np
import pytest
import torch

import acoustic_echo_cancellation_b200 as A
from acoustic_echo_cancellation_b200 import _lib, synth
from oracle import aec_oracle as O

pytestmark = pytest.mark.gpu

TOL_ERR = 1e-4      # max abs on the time-domain error signal (north star)
TOL_ERLE = 0.05     # dB


def _cuda(x):
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


def _run_case(P, algo, L, B=3, ragged=False, double_talk=False, variant=0, first_u=0, echo=True, frame=512):
    hop = frame // 2
    d = synth.make_batch(first_u, B, L, sample_rate=16000 * frame // 512, rir_len=min(P * hop, 4096),
                         double_talk=double_talk)
    ns = None
    if ragged:
        ns = np.array(([L, L - 1, max(L - 777, 1), 255, 0, 256, 257] * B)[:B], dtype=np.int64)
    skip = 8
    ref = O.stage1(d["far"], d["mic"], O.AecConfig(frame=frame, partitions=P, algo=algo, delta=1e-6 * frame),
                   n_samples=ns, erle_skip=skip * hop)
    cfg = A.Stage1Config(frame=frame, partitions=P, algo=algo, erle_skip_hops=skip, variant=variant)
    res = A.stage1_aec(_cuda(d["far"]), _cuda(d["mic"]), cfg, n_samples=None if ns is None else _cuda(ns),
                       return_echo=echo, return_erle=True)
    torch.cuda.synchronize()
    err, erle = res[0].cpu().numpy(), res[-1].cpu().numpy()
    n = ref["err"].shape[1]
    if n:
        assert np.abs(err[:, :n] - ref["err"]).max() <= TOL_ERR
    assert (err[:, n:] == 0).all()
    if echo:
        ec = res[1].cpu().numpy()
        if n:
            assert np.abs(ec[:, :n] - ref["echo"]).max() <= TOL_ERR
        assert (ec[:, n:] == 0).all()
    lens = ns if ns is not None else [L] * B
    for b in range(B):
        m = max(O.n_frames(int(lens[b]), frame, hop) - 1, 0) * hop
        assert (err[b, m:] == 0).all()
        if m > skip * hop:
            assert abs(erle[b] - ref["erle_db"][b]) <= TOL_ERLE
    return err, ref


@pytest.mark.parametrize("P,algo", [(4, 0), (4, 1), (1, 0), (1, 1), (2, 0), (2, 1), (8, 0), (8, 1), (16, 0), (16, 1)])
def test_stage1_matches_oracle(P, algo):
    _run_case(P, algo, 16000)


def test_config1_single_10s_utterance():
    """BASELINE.json configs[0]: single 10 s 16 kHz pair, frame 512, hop 256, 4 partitions."""
    err, ref = _run_case(4, 0, 160000, B=1)
    assert err.shape == (1, 160000) and ref["err"].shape[1] == 160000


def test_long_tail_kalman_10s():
    """configs[2] algorithm (16-partition Kalman) on 10 s utterances."""
    _run_case(16, 1, 160000, B=2)


@pytest.mark.parametrize("P,algo", [(8, 0), (8, 1), (4, 0), (4, 1), (2, 0), (1, 1)])
def test_frame_1024_matches_oracle(P, algo):
    """BASELINE.json configs[3] geometry: 48 kHz, frame 1024, hop 512 (8 partitions is the named case)."""
    _run_case(P, algo, 24000 + 77, ragged=True, B=4, frame=1024)


def test_config4_48khz_double_talk_10s():
    _run_case(8, 0, 480000, B=2, double_talk=True, frame=1024, echo=False)


@pytest.mark.parametrize("P,algo", [(8, 2), (8, 3), (4, 2), (4, 3)])
def test_frame_1024_overlap_save_matches_oracle(P, algo):
    """the overlap-save filters at the configs[3] geometry (blocks of 512 samples, FFT 1024): ragged lengths, echo estimate"""
    _run_case(P, algo, 24000 + 77, ragged=True, B=4, frame=1024)


def test_config4_48khz_double_talk_10s_overlap_save_kalman():
    """configs[3] as worded (48 kHz, frame 1024, 8 partitions, double-talk mixes) through algo 3, 10 s"""
    _run_case(8, 3, 480000, B=2, double_talk=True, frame=1024)


def test_double_talk_mixes():
    _run_case(4, 0, 32000, double_talk=True)
    _run_case(4, 1, 32000, double_talk=True)


@pytest.mark.parametrize("L", [16000 + 123, 4097, 300])
def test_ragged_and_degenerate_lengths(L):
    _run_case(4, 0, L, B=7, ragged=True)


@pytest.mark.parametrize("P,algo,variant", [(16, 1, 0), (16, 0, 0), (16, 0, 8128), (8, 1, 0), (8, 0, 0)])
@pytest.mark.parametrize("L", [16000 + 123, 8 * 256, 8 * 256 + 1, 9 * 256 - 1, 300])
def test_long_filters_ragged_and_chunk_boundaries(P, algo, variant, L):
    """The long-filter kernels (8 / 16 frames per chunk; for 16 partitions bin 128 runs one chunk ahead on the
    warps without synthesis work): lengths around the chunk boundary, ragged batches with 0- and 1-hop
    utterances, echo output on."""
    _run_case(P, algo, L, B=7, ragged=True, variant=variant)


def test_long_filter_unaligned_rows():
    """16-partition Kalman on rows that are not 16-byte aligned: every hop goes through the predicated-load
    path, which the look-ahead job of bin 128 reads as well."""
    L, B = 16001, 3
    d = synth.make_batch(0, B, L, rir_len=4096)
    ref = O.stage1(d["far"], d["mic"], O.AecConfig(partitions=16, algo=O.ALGO_KALMAN))
    bf = torch.zeros(B, L + 3, device="cuda")
    bm = torch.zeros(B, L + 3, device="cuda")
    bo = torch.zeros(B, L + 3, device="cuda")
    bf[:, 1:L + 1] = _cuda(d["far"])
    bm[:, 1:L + 1] = _cuda(d["mic"])
    cfg = A.Stage1Config(partitions=16, algo=A.ALGO_KALMAN)
    err = A.stage1_aec(bf[:, 1:L + 1], bm[:, 1:L + 1], cfg, out=bo[:, 1:L + 1]).cpu().numpy()
    n = ref["err"].shape[1]
    assert np.abs(err[:, :n] - ref["err"]).max() <= TOL_ERR
    assert float(bo[:, 0].abs().max()) == 0 and float(bo[:, L + 1:].abs().max()) == 0   # no stray writes


@pytest.mark.parametrize("variant", [2128, 2168, 4128, 4096, 1255, 1200])
def test_tuning_variants_agree(variant):
    _run_case(4, 0, 16000 + 256, variant=variant, echo=False)   # tuning variants are built without the echo output


def test_unaligned_rows_take_the_in_kernel_slow_path():
    L, B = 16001, 3
    d = synth.make_batch(0, B, L)
    ref = O.stage1(d["far"], d["mic"], O.AecConfig())
    bf = torch.zeros(B, L + 3, device="cuda")
    bm = torch.zeros(B, L + 3, device="cuda")
    bo = torch.zeros(B, L + 3, device="cuda")
    bf[:, 1:L + 1] = _cuda(d["far"])
    bm[:, 1:L + 1] = _cuda(d["mic"])
    err = A.stage1_aec(bf[:, 1:L + 1], bm[:, 1:L + 1], out=bo[:, 1:L + 1]).cpu().numpy()
    n = ref["err"].shape[1]
    assert np.abs(err[:, :n] - ref["err"]).max() <= TOL_ERR
    assert float(bo[:, 0].abs().max()) == 0 and float(bo[:, L + 1:].abs().max()) == 0   # no stray writes


@pytest.mark.parametrize("frame,P,algo,echo", [(512, 4, 0, False), (512, 4, 1, True), (512, 8, 1, False),
                                               (512, 16, 1, True), (512, 16, 0, False), (1024, 8, 0, True)])
def test_no_stray_writes_and_run_to_run_determinism(frame, P, algo, echo):
    """compute-sanitizer is closed on this pool: canaries around every output row and bitwise
    run-to-run equality (a data race would show as a flaky difference) stand in for it."""
    B, L, pad = 5, 9000, 64
    hop = frame // 2
    g = torch.Generator(device="cuda").manual_seed(3)
    far = 0.1 * torch.randn(B, L, device="cuda", generator=g)
    mic = 0.5 * torch.roll(far, 11, dims=1) + 0.01 * torch.randn(B, L, device="cuda", generator=g)
    ns = torch.tensor([L, L - 1, L - hop - 3, 17, L // 2], device="cuda")
    cfg = A.Stage1Config(frame=frame, partitions=P, algo=algo, erle_skip_hops=2)
    canary = 12345.0
    stride = L + 2 * pad
    stride += (-stride) % 4
    outs = []
    for _ in range(2):
        buf = torch.full((B, stride), canary, device="cuda")
        ebuf = torch.full((B, stride), canary, device="cuda")
        lib = _lib.load()
        c = cfg.to_c()
        erle = torch.full((B + 2,), canary, device="cuda")
        rc = lib.aec_stage1_run(far.data_ptr(), mic.data_ptr(), buf[:, pad:].data_ptr(),
                                ebuf[:, pad:].data_ptr() if echo else None, erle[1:].data_ptr(), ns.data_ptr(),
                                B, L, L, stride, C.byref(c), int(torch.cuda.current_stream().cuda_stream))
        assert rc == 0
        torch.cuda.synchronize()
        for t in ((buf, ebuf) if echo else (buf,)):
            assert bool((t[:, :pad] == canary).all()) and bool((t[:, pad + L:] == canary).all())
            assert bool(torch.isfinite(t[:, pad:pad + L]).all())
        assert float(erle[0]) == canary and float(erle[-1]) == canary
        outs.append((buf.clone(), ebuf.clone(), erle.clone()))
    for a, b in zip(outs[0], outs[1]):
        assert torch.equal(a, b)


@pytest.mark.parametrize("P,algo,B", [(4, 0, 1100), (16, 1, 320), (16, 0, 320), (8, 1, 320)])
def test_full_occupancy_runs_are_bitwise_repeatable(P, algo, B):
    """Every SM fully occupied (more than one wave), ragged lengths, five launches: any ordering bug between the warps
    of an utterance (named barrier of the look-ahead job, bulk-copy runs, overlap-add hand-over) would show up as a
    run-to-run difference under this much contention.  compute-sanitizer is not available on the GPU pool."""
    L = 256 * 41 + 77
    g = torch.Generator(device="cuda").manual_seed(11)
    far = 0.1 * torch.randn(B, L, device="cuda", generator=g)
    mic = 0.5 * torch.roll(far, 19, dims=1) + 0.01 * torch.randn(B, L, device="cuda", generator=g)
    ns = (L - (torch.arange(B, device="cuda") * 37) % (L // 2)).to(torch.int64)
    cfg = A.Stage1Config(partitions=P, algo=algo, erle_skip_hops=2)
    ref = None
    for _ in range(5):
        err, erle = A.stage1_aec(far, mic, cfg, n_samples=ns, return_erle=True)
        torch.cuda.synchronize()
        assert bool(torch.isfinite(err).all())
        if ref is None:
            ref = (err.clone(), erle.clone())
        else:
            assert torch.equal(err, ref[0]) and torch.equal(erle, ref[1])
    # and utterance b is what it is when processed alone (batch invariance under full occupancy)
    b = B // 2
    alone = A.stage1_aec(far[b:b + 1], mic[b:b + 1], cfg, n_samples=ns[b:b + 1])
    assert torch.equal(alone[0], ref[0][b])


def test_wav2h5_runner_pcm16_and_tapered_slices_match_the_device_path():
    """The generator's runner (host pipeline, echo output on) on a batch that exercises the tapered last slices
    (300 utterances in slices of 64: 64 x 4, 22, 11, 11) gives the device path's results, and the int16 PCM form of the
    same signals (what create_h5 sends for a 16-bit wav corpus) gives bit-identical ones."""
    from acoustic_echo_cancellation_b200 import wav2h5
    B, L = 300, 2048 + 100
    rng = np.random.default_rng(3)
    pcm_f = rng.integers(-6000, 6000, size=(B, L), dtype=np.int16)
    pcm_m = np.roll(pcm_f, 7, axis=1) // 2
    far, mic = pcm_f.astype(np.float32) / 32768.0, pcm_m.astype(np.float32) / 32768.0
    n = np.full(B, L, dtype=np.int64)
    n[5], n[77] = 900, 257
    run = wav2h5.default_runner(slice_utterances=64)
    err_f, echo_f = run(far, mic, n)
    err_p, echo_p = run(pcm_f, pcm_m, n)
    assert err_f.dtype == np.float32 and err_p.dtype == np.float32
    assert np.array_equal(err_f, err_p) and np.array_equal(echo_f, echo_p)
    dev = A.stage1_aec(_cuda(far), _cuda(mic), A.Stage1Config(), n_samples=_cuda(n), return_echo=True)
    assert np.array_equal(err_f, dev[0].cpu().numpy()) and np.array_equal(echo_f, dev[1].cpu().numpy())


@pytest.mark.parametrize("algo,P", [(0, 4), (3, 4), (2, 8)])
def test_host_buffer_entry_matches_device_entry(algo, P):
    L, B = 16000, 10
    d = synth.make_batch(0, B, L)
    cfg = A.Stage1Config(erle_skip_hops=4, algo=algo, partitions=P)
    dev, dev_erle = A.stage1_aec(_cuda(d["far"]), _cuda(d["mic"]), cfg, return_erle=True)
    pipe = A.HostPipeline(slice_utterances=4, max_samples=L)          # 3 slices, last one partial
    err = np.empty((B, L), dtype=np.float32)
    echo = np.empty((B, L), dtype=np.float32)
    erle = np.empty(B, dtype=np.float32)
    pipe.run(d["far"], d["mic"], cfg, err=err, echo=echo, erle=erle)
    pipe.close()
    assert np.array_equal(err, dev.cpu().numpy())
    assert np.array_equal(erle, dev_erle.cpu().numpy())
    n = A.out_samples(L)
    assert np.abs(err + echo - d["mic"])[:, :n].max() < 1e-5


def test_pcm16_host_entry_equals_float_entry_on_quantised_input():
    L, B = 16000, 6
    d = synth.make_batch(0, B, L)
    far16 = np.round(d["far"] * 32768).clip(-32768, 32767).astype(np.int16)
    mic16 = np.round(d["mic"] * 32768).clip(-32768, 32767).astype(np.int16)
    pipe = A.HostPipeline(slice_utterances=4, max_samples=L)
    e16 = pipe.run(far16, mic16)
    ef = pipe.run(far16.astype(np.float32) / 32768, mic16.astype(np.float32) / 32768)
    pipe.close()
    assert np.array_equal(e16, ef)


def test_host_pipeline_geometry_does_not_change_the_result():
    """slots in flight, slice size and the start ramp are scheduling only: every geometry gives the device path's bits"""
    B, L = 200, 4096 + 50
    d = synth.make_batch(40, 8, L)
    far, mic = np.tile(d["far"], (B // 8, 1)), np.tile(d["mic"], (B // 8, 1))
    far[3] *= 0.5
    cfg = A.Stage1Config(erle_skip_hops=2)
    dev, dev_erle = A.stage1_aec(_cuda(far), _cuda(mic), cfg, return_erle=True)
    dev, dev_erle = dev.cpu().numpy(), dev_erle.cpu().numpy()
    hf, hm = A.pinned_empty((B, L)), A.pinned_empty((B, L))
    hf[:], hm[:] = far, mic
    assert A.is_pinned(hf) and not A.is_pinned(far)
    for slots, sl, ramp in [(1, 16, True), (2, 32, False), (4, 16, True), (8, 8, True), (3, 64, True)]:
        pipe = A.HostPipeline(sl, L, slots=slots, ramp=ramp)
        err = A.pinned_empty((B, L))
        erle = np.zeros(B, dtype=np.float32)
        pipe.run(hf, hm, cfg, err=err, erle=erle)
        pipe.close()
        assert np.array_equal(err, dev) and np.array_equal(erle, dev_erle), (slots, sl, ramp)
    with pytest.raises(A.AecError):
        A.HostPipeline(16, L, slots=9)
    # streaming (deferred) mode: calls return before their last slices have landed; after wait() every batch is the
    # device path's result, whatever was in flight when the next call started
    pipe = A.HostPipeline(32, L, slots=4)
    outs = [A.pinned_empty((B, L)) for _ in range(3)]
    erles = [np.zeros(B, dtype=np.float32) for _ in range(3)]
    hf2 = A.pinned_empty((B, L))
    hf2[:] = far[::-1]
    hm2 = A.pinned_empty((B, L))
    hm2[:] = mic[::-1]
    pipe.run(hf, hm, cfg, err=outs[0], erle=erles[0], wait=False)
    pipe.run(hf2, hm2, cfg, err=outs[1], erle=erles[1], wait=False)
    pipe.run(hf, hm, cfg, err=outs[2], erle=erles[2], wait=False)
    pipe.wait()
    assert np.array_equal(outs[0], dev) and np.array_equal(outs[2], dev) and np.array_equal(outs[1], dev[::-1])
    assert np.array_equal(erles[0], dev_erle) and np.array_equal(erles[1], dev_erle[::-1]) and np.array_equal(erles[2], dev_erle)
    e4 = pipe.run(hf, hm, cfg, erle=erles[0])                     # back to the synchronous form on the same context
    assert np.array_equal(e4, dev)
    pipe.close()


def test_runner_buffers_are_page_locked_and_pageable_inputs_are_staged():
    """the generator's runner must not hand pageable arrays to cudaMemcpyAsync"""
    from acoustic_echo_cancellation_b200 import wav2h5
    B, L = 40, 3000
    d = synth.make_batch(7, B, L)
    run = wav2h5.default_runner(slice_utterances=16)
    n = np.full(B, L, dtype=np.int64)
    err, echo = run(d["far"], d["mic"], n)                       # pageable numpy in
    assert run.buffers_pinned() and A.is_pinned(err) and A.is_pinned(echo)
    assert set(run._in) == {"far", "mic"}                         # ... so they were staged
    dev = A.stage1_aec(_cuda(d["far"]), _cuda(d["mic"]), A.Stage1Config(), return_echo=True)
    assert np.array_equal(err, dev[0].cpu().numpy()) and np.array_equal(echo, dev[1].cpu().numpy())
    err1 = err.copy()
    err2, _ = run(d["far"][::-1].copy(), d["mic"][::-1].copy(), n)   # next call uses the OTHER output set
    assert np.array_equal(err, err1) and np.array_equal(err2[::-1], err1)
    run.close()


class _ExFile:
    """`z[key]` -> array for an .ex (HDF5) file: h5py when present, else h5lite's reader"""

    def __init__(self, path):
        try:
            import h5py  # type: ignore
            self.f = h5py.File(path, "r")
        except ImportError:
            from acoustic_echo_cancellation_b200 import h5lite
            self.f = h5lite.File(path, "r")

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.f.close()

    def __getitem__(self, key):
        return np.array(self.f[key])


@pytest.mark.parametrize("algo_flag", [None, "ols-kalman"])
def test_create_h5_train_end_to_end_on_the_gpu(tmp_path, algo_flag):
    """wav files -> create_h5 (decode | stage 1 | write, default CUDA runner, NpzStore container) -> files whose
    stage-1 datasets are the device path's output on the decoded signals; with --stage1_algo ols-kalman the runner the
    command-line flags build (wav2h5.runner_from_args) puts the overlap-save filter with the Kalman step there"""
    import types

    from scipy.io import wavfile

    from acoustic_echo_cancellation_b200 import wav2h5
    wav_dir, h5_dir, list_dir = tmp_path / "wav", tmp_path / "h5", tmp_path / "lists"
    for p in (wav_dir, h5_dir, list_dir):
        p.mkdir()
    ids = [str(i) for i in range(37)]
    sig = {}
    for i in ids:
        u = synth.make_utterance(int(i), 5000 + 64 * int(i), rir_len=512)
        for key, src in (("farend_speech", "far"), ("nearend_mic", "mic"), ("echo", "echo"), ("nearend_speech", "near")):
            pcm = np.clip(np.rint(u[src] * 32768.0), -32768, 32767).astype(np.int16)
            wavfile.write(str(wav_dir / wav2h5.WAV_PATTERNS[key].format(idx=i)), 16000, pcm)
            sig[(i, key)] = pcm.astype(np.float32) / np.float32(32768)
    args = types.SimpleNamespace(train_path=str(wav_dir), h5_path=str(h5_dir), list_path=str(list_dir), sr=16000)
    st = {}
    cfg = A.Stage1Config()
    kw = {}
    if algo_flag:
        a = wav2h5.build_parser("train").parse_args(["--train_path", str(wav_dir), "--h5_path", str(h5_dir), "--list_path",
                                                     str(list_dir), "--stage1_algo", algo_flag, "--stage1_partitions", "4"])
        kw["runner"] = wav2h5.runner_from_args(a)
        cfg = A.Stage1Config(algo=A.ALGO_PBFKF, partitions=4)
    # the plain run writes the stand-in npz container, the ols-kalman run the DEFAULT one: real HDF5 (.ex) through
    # h5py when installed, else through the package's own writer (h5lite)
    from acoustic_echo_cancellation_b200 import h5lite
    if algo_flag:
        paths = wav2h5.create_h5(args, batch=8, stats=st, **kw)
        assert all(h5lite.is_hdf5(p) for p in paths)
    else:
        paths = wav2h5.create_h5(args, batch=8, h5=wav2h5.NpzStore(), stats=st, **kw)
    assert len(paths) == 37 and st["pcm16_batches"] == 5 and st["float32_batches"] == 0
    for p in paths[::6]:
        i = p.split("tr_")[-1][:-3]
        with (_ExFile(p) if algo_flag else np.load(p)) as z:
            for key in wav2h5.KEYS:
                assert np.array_equal(z[key], sig[(i, key)])
            dev = A.stage1_aec(_cuda(sig[(i, "farend_speech")][None]), _cuda(sig[(i, "nearend_mic")][None]), cfg,
                               return_echo=True)
            assert np.array_equal(z["stage1_error"], dev[0].cpu().numpy()[0])
            assert np.array_equal(z["stage1_echo"], dev[1].cpu().numpy()[0])


def test_full_batch_against_the_c_port_on_random_utterances():
    """at the full config-2 and config-3 batch sizes, 32 randomly chosen utterances of the batch
    against the float32 C port (fast enough for 10 s utterances) instead of relying on batch-invariance transitivity;
    the config-2 batch also through the overlap-save filters (algos 2 / 3; the port runs them eight utterances abreast)"""
    from oracle import c_oracle as CO
    rng = np.random.default_rng(11)
    for (B, P, algo, L) in [(1024, 4, 0, 160000), (4096, 16, 1, 160000), (1024, 4, 2, 160000), (1024, 4, 3, 160000)]:
        base = synth.make_batch(500, 16, L, rir_len=min(P * 256, 2048))
        gains = (0.25 + 0.75 * rng.random(B)).astype(np.float32)
        far = torch.from_numpy(base["far"]).cuda()[torch.arange(B, device="cuda") % 16] * torch.from_numpy(gains).cuda()[:, None]
        mic = torch.from_numpy(base["mic"]).cuda()[torch.arange(B, device="cuda") % 16] * torch.from_numpy(gains).cuda()[:, None]
        cfg = A.Stage1Config(partitions=P, algo=algo, erle_skip_hops=125)
        err, erle = A.stage1_aec(far, mic, cfg, return_erle=True)
        pick = np.sort(rng.choice(B, size=32, replace=False))
        hf, hm = far[pick].cpu().numpy(), mic[pick].cpu().numpy()
        ref = CO.stage1(hf, hm, O.AecConfig(partitions=P, algo=algo), want_echo=False, erle_skip_hops=125)
        got = err[pick].cpu().numpy()
        n = A.out_samples(L)
        assert np.abs(got[:, :n] - ref["err"][:, :n]).max() <= TOL_ERR, (B, P, algo)
        assert np.abs(erle[pick].cpu().numpy() - ref["erle_db"]).max() <= TOL_ERLE
        del far, mic, err
        torch.cuda.empty_cache()


@pytest.mark.parametrize("P,algo,L,B", [(4, 0, 16000 + 123, 5), (4, 1, 8 * 256, 3), (2, 0, 4097, 2), (1, 0, 300, 2),
                                       (4, 0, 160000, 40)])
def test_fused_feature_epilogue_equals_features_of_the_stored_error(P, algo, L, B):
    """stage 1 with the Stage-2 front end fused in (aec_stage1_run_features) gives (a) the same error
    signal, bit for bit, as the plain kernel and (b) the features the stand-alone operator computes from that stored
    error signal and the far end (ERB.py:262-290, in_norm off) -- the parity-preserving form: STFT of the SYNTHESISED
    error, not the filter's internal spectrum."""
    d = synth.make_batch(30, B, L, rir_len=min(P * 256, 1024))
    far, mic = _cuda(d["far"]), _cuda(d["mic"])
    erb = torch.from_numpy(A.erb_filterbank()).float().cuda()
    cfg = A.Stage1Config(partitions=P, algo=algo, erle_skip_hops=2)
    err, feat, erle = A.stage1_aec_features(far, mic, erb, cfg, return_erle=True)
    ref_err, ref_erle = A.stage1_aec(far, mic, cfg, return_erle=True)
    assert torch.equal(err, ref_err) and torch.equal(erle, ref_erle)
    want = A.stage2_features(ref_err, far, erb, in_norm=False)
    assert feat.shape == want.shape == (B, A.num_frames(L), 64)
    scale = float(want.abs().max())
    assert float((feat - want).abs().max()) <= 2e-5 * max(scale, 1.0), float((feat - want).abs().max())
    assert torch.equal(feat, A.stage1_aec_features(far, mic, erb, cfg)[1])          # run-to-run determinism
    # the filter's own spectrum would NOT do: features of the microphone signal differ from those of the error
    assert float((feat - A.stage2_features(mic, far, erb, in_norm=False)).abs().max()) > 1e-3 * scale or L < 1000


def test_fused_feature_epilogue_ragged_rows_and_golden_front_end(golden):
    """ragged batch: rows of an utterance are those of the utterance processed alone, rows beyond its own frames hold
    the front end's response to silence; and with a zero far end (error = microphone) the fused features of
    (feat_mic, 0) equal the reference-generated golden front end of (feat_mic, 0)-shaped inputs computed stand-alone"""
    L, B = 6000, 4
    d = synth.make_batch(60, B, L, rir_len=512)
    far, mic = _cuda(d["far"]), _cuda(d["mic"])
    erb = torch.from_numpy(A.erb_filterbank()).float().cuda()
    ns = torch.tensor([L, 3000, 255, 4096], dtype=torch.int64, device="cuda")
    cfg = A.Stage1Config()
    err, feat = A.stage1_aec_features(far, mic, erb, cfg, n_samples=ns)
    assert torch.equal(err, A.stage1_aec(far, mic, cfg, n_samples=ns))
    silent = A.stage2_features(torch.zeros(1, 512, device="cuda"), torch.zeros(1, 512, device="cuda"), erb, in_norm=False)[0, 0]
    for b in range(B):
        n = int(ns[b])
        tb = A.num_frames(n)
        e1, f1 = A.stage1_aec_features(far[b:b + 1, :n].contiguous(), mic[b:b + 1, :n].contiguous(), erb, cfg)
        assert torch.equal(feat[b, :tb], f1[0])
        assert torch.allclose(feat[b, tb:], silent.expand(feat.shape[1] - tb, 64), rtol=0, atol=1e-9)
    # zero far end: the canceller is the identity (to float rounding), so the fused features are the front end of the mic
    gm = _cuda(golden["feat_mic"])
    e, f = A.stage1_aec_features(torch.zeros_like(gm), gm, erb, cfg)
    want = A.stage2_features(e, torch.zeros_like(gm), erb, in_norm=False)
    assert float((f - want).abs().max()) <= 2e-5 * max(float(want.abs().max()), 1.0)


@pytest.mark.parametrize("B,P,algo,frame", [(1024, 4, 2, 512), (1024, 4, 3, 512), (2048, 16, 3, 512), (512, 8, 3, 1024)])
def test_overlap_save_full_batches_against_the_oracle_on_random_utterances(B, P, algo, frame):
    """full-occupancy batches (every SM holds its maximum of co-resident utterances) through the overlap-save kernels: 12
    randomly chosen utterances of the batch against the float64 numpy oracle, 10 s each"""
    rng = np.random.default_rng(17)
    hop = frame // 2
    L = 160000 * frame // 512
    base = synth.make_batch(600, 8, L, sample_rate=16000 * frame // 512, rir_len=min(P * hop, 4096))
    gains = (0.25 + 0.75 * rng.random(B)).astype(np.float32)
    idx = torch.arange(B, device="cuda") % 8
    far = torch.from_numpy(base["far"]).cuda()[idx] * torch.from_numpy(gains).cuda()[:, None]
    mic = torch.from_numpy(base["mic"]).cuda()[idx] * torch.from_numpy(gains).cuda()[:, None]
    cfg = A.Stage1Config(frame=frame, partitions=P, algo=algo, erle_skip_hops=125)
    err, erle = A.stage1_aec(far, mic, cfg, return_erle=True)
    pick = np.sort(rng.choice(B, size=12, replace=False))
    ref = O.stage1(far[pick].cpu().numpy(), mic[pick].cpu().numpy(),
                   O.AecConfig(frame=frame, partitions=P, algo=algo, delta=1e-6 * frame), erle_skip=125 * hop)
    n = ref["err"].shape[1]
    assert np.abs(err[pick].cpu().numpy()[:, :n] - ref["err"]).max() <= TOL_ERR
    assert np.abs(erle[pick].cpu().numpy() - ref["erle_db"]).max() <= TOL_ERLE


@pytest.mark.parametrize("algo", [2, 3])
@pytest.mark.parametrize("P,L,B", [(4, 16000 + 123, 3), (2, 8 * 256, 2), (1, 4097, 2), (4, 300, 2), (4, 160000, 2),
                                   (8, 16000 + 123, 3), (16, 24000 + 5, 3), (8, 160000, 1), (16, 160000, 1)])
def test_overlap_save_pbfdaf_matches_oracle(P, L, B, algo):
    """algo = 2 / 3 (overlap-save PBFDAF, alternated constraint, NLMS / Kalman step; builder-authored, parity unpinned):
    error signal, echo estimate and ERLE against the float64 numpy oracle; same tolerance as the STFT-domain recurrences"""
    d = synth.make_batch(70, B, L, rir_len=P * 256)
    ns = np.array(([L, max(L - 777, 1), 255] * B)[:B], dtype=np.int64) if L < 100000 else None
    skip = 8
    ref = O.stage1(d["far"], d["mic"], O.AecConfig(partitions=P, algo=algo), n_samples=ns, erle_skip=skip * 256)
    cfg = A.Stage1Config(partitions=P, algo=algo, erle_skip_hops=skip)
    err, echo, erle = A.stage1_aec(_cuda(d["far"]), _cuda(d["mic"]), cfg, n_samples=None if ns is None else _cuda(ns),
                                   return_echo=True, return_erle=True)
    torch.cuda.synchronize()
    err, echo, erle = err.cpu().numpy(), echo.cpu().numpy(), erle.cpu().numpy()
    n = ref["err"].shape[1]
    if n:
        assert np.abs(err[:, :n] - ref["err"]).max() <= TOL_ERR
        assert np.abs(echo[:, :n] - ref["echo"]).max() <= TOL_ERR
    assert (err[:, n:] == 0).all() and (echo[:, n:] == 0).all()
    lens = ns if ns is not None else [L] * B
    for b in range(B):
        m = (int(lens[b]) // 256) * 256
        assert (err[b, m:] == 0).all()
        if m > skip * 256:
            assert abs(erle[b] - ref["erle_db"][b]) <= TOL_ERLE
    # no echo output requested: same error signal, bit for bit; run-to-run determinism
    e2 = A.stage1_aec(_cuda(d["far"]), _cuda(d["mic"]), cfg, n_samples=None if ns is None else _cuda(ns))
    assert np.array_equal(e2.cpu().numpy(), err)


@pytest.mark.parametrize("P", [4, 8, 16])
@pytest.mark.parametrize("algo", [2, 3])
def test_overlap_save_unaligned_rows_and_batch_invariance(algo, P):
    """rows that are not 16-byte aligned take the synchronous staging path; an utterance's result does not depend on
    what else is in the batch"""
    L = 20 * 256 + 77
    d = synth.make_batch(90, 5, L, rir_len=1024)
    cfg = A.Stage1Config(partitions=P, algo=algo)
    far, mic = _cuda(d["far"]), _cuda(d["mic"])
    base = A.stage1_aec(far, mic, cfg)
    pad_f = torch.zeros(5, L + 3, device="cuda")
    pad_m = torch.zeros(5, L + 3, device="cuda")
    pad_f[:, 1:L + 1] = far
    pad_m[:, 1:L + 1] = mic
    odd = A.stage1_aec(pad_f[:, 1:L + 1], pad_m[:, 1:L + 1], cfg)
    assert torch.equal(odd, base)
    one = A.stage1_aec(far[3:4].contiguous(), mic[3:4].contiguous(), cfg)
    assert torch.equal(one[0], base[3])


@pytest.mark.parametrize("algo,P,frame", [(2, 4, 512), (3, 4, 512), (3, 16, 512), (3, 8, 1024)])
def test_overlap_save_oracle_free_properties(algo, P, frame):
    """silent far end: the canceller is the identity, bit for bit (y = IFFT(0)); digital silence everywhere: zeros and a
    finite ERLE (the regularisers); e + y = d to rounding; a 60 s utterance stays finite and converged"""
    sr = 16000 * frame // 512
    hop = frame // 2
    d = synth.make_batch(40, 3, 6 * sr + 37, sample_rate=sr, rir_len=min(P * hop, 4096))
    cfg = A.Stage1Config(frame=frame, partitions=P, algo=algo, erle_skip_hops=4)
    mic = _cuda(d["mic"])
    n = (mic.shape[1] // hop) * hop
    err, echo = A.stage1_aec(torch.zeros_like(mic), mic, cfg, return_echo=True)
    assert torch.equal(err[:, :n], mic[:, :n]) and not bool(echo.any())
    err, echo, erle = A.stage1_aec(torch.zeros_like(mic), torch.zeros_like(mic), cfg, return_echo=True, return_erle=True)
    assert not bool(err.any()) and not bool(echo.any()) and bool(torch.isfinite(erle).all())
    err, echo = A.stage1_aec(_cuda(d["far"]), mic, cfg, return_echo=True)
    assert float((err + echo - mic)[:, :n].abs().max()) <= 1e-6
    long = synth.make_batch(41, 1, 60 * sr, sample_rate=sr, rir_len=min(P * hop, 4096))
    e60, erle60 = A.stage1_aec(_cuda(long["far"]), _cuda(long["mic"]), A.Stage1Config(frame=frame, partitions=P, algo=algo,
                                                                                     erle_skip_hops=30 * sr // hop),
                               return_erle=True)
    assert bool(torch.isfinite(e60).all()) and float(erle60[0]) > 25.0


def test_overlap_save_pbfdaf_reaches_the_noise_floor_where_the_stft_recurrence_does_not():
    """the reason algos 2 / 3 exist: single talk, 4 partitions, -40 dB noise"""
    d = synth.make_batch(0, 4, 160000, rir_len=1024)
    far, mic = _cuda(d["far"]), _cuda(d["mic"])
    _, erle_ols = A.stage1_aec(far, mic, A.Stage1Config(algo=A.ALGO_PBFDAF, erle_skip_hops=250), return_erle=True)
    _, erle_kf = A.stage1_aec(far, mic, A.Stage1Config(algo=A.ALGO_PBFKF, erle_skip_hops=250), return_erle=True)
    _, erle_stft = A.stage1_aec(far, mic, A.Stage1Config(algo=A.ALGO_NLMS, erle_skip_hops=250), return_erle=True)
    assert float(erle_ols.min()) > 35.0 and float(erle_kf.min()) > 30.0 and float(erle_stft.max()) < 20.0


def test_overlap_save_kalman_step_holds_through_double_talk():
    """algo 3 vs algo 2 on the SURVEY 8d double-talk set: echo-only ERLE (echo vs echo - echo estimate) over the
    double-talk span [0.4 L, 0.7 L]; the NLMS step is pulled away by the near end, the Kalman step is not"""
    L = 160000
    d = synth.make_batch(0, 4, L, rir_len=1024, double_talk=True)
    far, mic = _cuda(d["far"]), _cuda(d["mic"])
    a, b = int(0.4 * L), int(0.7 * L)

    def dt_erle(algo):
        _, echo = A.stage1_aec(far, mic, A.Stage1Config(algo=algo), return_echo=True)
        res = d["echo"][:, a:b] - echo.cpu().numpy()[:, a:b]
        return 10 * np.log10((d["echo"][:, a:b] ** 2).sum(1) / (res ** 2).sum(1))

    nlms, kal = dt_erle(A.ALGO_PBFDAF), dt_erle(A.ALGO_PBFKF)
    assert kal.min() > 10.0 and nlms.max() < 8.0 and (kal - nlms).min() > 5.0


def test_batch_shift_matches_torch_and_stays_on_the_device():
    """ERB.py:254-256: mean / std (unbiased) over the whole batch tensor, reduced on the device (no host sync)"""
    g = torch.Generator(device="cuda").manual_seed(3)
    for shape, off in (((7, 4099), 0.01), ((64, 160000), -0.003), ((1, 513), 0.2)):
        x = 0.1 * torch.randn(shape, device="cuda", generator=g) + off
        s = A.batch_shift(x)
        assert s.is_cuda and s.shape == (1,)
        want = (x.double().mean() / x.double().std()).item()
        assert abs(s.item() - want) <= 2e-6 * max(1.0, abs(want))
        assert torch.equal(s, A.batch_shift(x))                       # fixed summation order: bitwise repeatable
    xs = torch.empty(5, 5000, device="cuda")[:, :4097]                # strided rows, odd length -> scalar path
    xs.copy_(0.05 * torch.randn(5, 4097, device="cuda", generator=g) + 0.02)
    assert abs(A.batch_shift(xs).item() - (xs.double().mean() / xs.double().std()).item()) <= 1e-6


def test_unsupported_combination_is_reported_not_emulated():
    d = synth.make_batch(0, 1, 4096)
    with pytest.raises(A.AecError) as ei:
        A.stage1_aec(_cuda(d["far"]), _cuda(d["mic"]), A.Stage1Config(partitions=5))
    assert ei.value.code == -2
    for cfg in (A.Stage1Config(partitions=3, algo=A.ALGO_PBFDAF),            # overlap-save: powers of two up to 16
                A.Stage1Config(partitions=16, algo=A.ALGO_PBFKF, frame=1024)):  # ... frame 1024: 4 or 8 partitions
        with pytest.raises(A.AecError) as ei:
            A.stage1_aec(_cuda(d["far"]), _cuda(d["mic"]), cfg)
        assert ei.value.code == -2
    with pytest.raises(A.AecError) as ei:                                       # no fused feature epilogue for them
        A.stage1_aec_features(_cuda(d["far"]), _cuda(d["mic"]), torch.from_numpy(A.erb_filterbank()).float().cuda(),
                              A.Stage1Config(algo=A.ALGO_PBFDAF))
    assert ei.value.code == -2


# ---------------------------------------------------------------------------------------------
# reference-pinned operators against the golden vectors
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("i", range(5))
def test_stft_istft_match_reference_golden(golden, i):
    x, s_ref, y_ref = golden[f"x512_{i}"], golden[f"stft512_{i}"], golden[f"istft512_{i}"]
    s = A.ConvSTFT(512, 256, 512, "hann", "complex")(_cuda(x))
    assert tuple(s.shape) == s_ref.shape
    assert np.abs(s.cpu().numpy() - s_ref).max() < 5e-5       # the reference's own conv path is this far from an FFT
    y = A.ConviSTFT(512, 256, 512, "hann", "complex")(_cuda(s_ref))
    assert tuple(y.shape) == y_ref.shape
    if y_ref.size:
        assert np.abs(y.cpu().numpy() - y_ref).max() < 5e-6


def test_stft_istft_frame_1024_golden(golden):
    s = A.ConvSTFT(1024, 512, 1024, "hann", "complex")(_cuda(golden["x1024"]))
    assert tuple(s.shape) == golden["stft1024"].shape
    assert np.abs(s.cpu().numpy() - golden["stft1024"]).max() < 1e-4
    y = A.ConviSTFT(1024, 512, 1024, "hann", "complex")(_cuda(golden["stft1024"]))
    assert tuple(y.shape) == golden["istft1024"].shape
    assert np.abs(y.cpu().numpy() - golden["istft1024"]).max() < 5e-6


def test_istft_free_spectrum_golden(golden):
    y = A.ConviSTFT(512, 256, 512, "hann", "complex")(_cuda(golden["spec_free"])).cpu().numpy()
    assert np.abs(y - golden["istft_free"]).max() < 5e-6


def test_feature_front_end_golden(golden):
    f = A.stage2_features(_cuda(golden["feat_mic"]), _cuda(golden["feat_ref"]),
                          _cuda(golden["erb"].astype(np.float32))).cpu().numpy()
    assert f.shape == golden["feat"].shape
    assert np.abs(f - golden["feat"]).max() < 2e-4 * np.abs(golden["feat"]).max()


def test_stage2_little_net_inference_matches_reference_module():
    """Stage-2 inference (ERB.py:252-316) against the output of the reference module itself."""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_stage2.npz"))
    w = {k[2:]: g[k] for k in g.files if k.startswith("w_")}
    net = A.LittleNetInference(w, g["erb"])
    out = net(_cuda(g["mic"]), _cuda(g["ref"])).cpu().numpy()
    assert out.shape == g["out_wav"].shape
    assert np.abs(out - g["out_wav"]).max() < 2e-4 * np.abs(g["out_wav"]).max()
    ref64 = O.stage2_little_net(g["mic"], g["ref"], g["erb"], w)
    assert np.abs(out - ref64).max() < 2e-4 * np.abs(ref64).max()


def test_batches_larger_than_the_grid_y_limit():
    """More than 65535 utterances per call (grid.y of the operator kernels carries the utterance index; the stage-1
    kernel uses grid.x): results must equal the same rows processed in a small batch."""
    B, L = 65535 + 700, 1024
    g = torch.Generator(device="cuda").manual_seed(5)
    x = 0.2 * torch.randn(B, L, device="cuda", generator=g)
    stft, istft = A.ConvSTFT(512, 256, 512, "hann", "complex"), A.ConviSTFT(512, 256, 512, "hann", "complex")
    s = stft(x)
    y = istft(s)
    rows = [0, 65534, 65535, B - 1]
    xs = x[rows].contiguous()
    assert torch.equal(s[rows], stft(xs)) and torch.equal(y[rows], istft(stft(xs)))
    erb = torch.from_numpy(O.erb_filterbank()).float().cuda()
    f = A.stage2_features(x, torch.roll(x, 5, dims=1), erb, in_norm=False)
    assert torch.equal(f[rows], A.stage2_features(xs, torch.roll(x, 5, dims=1)[rows].contiguous(), erb, in_norm=False))
    err = A.stage1_aec(torch.roll(x, 3, dims=1), x, A.Stage1Config(partitions=2))
    assert torch.equal(err[rows], A.stage1_aec(torch.roll(x, 3, dims=1)[rows].contiguous(), xs, A.Stage1Config(partitions=2)))


def test_full_size_operators_against_reference_summaries():
    """10 s utterances (BASELINE.json's length) through the stand-alone kernels and the Stage-2 inference kernels,
    against 256 pinned values and per-channel / per-frame sums of the REFERENCE's own outputs
    (tests/golden/reference_full_size.npz, written by tests/golden/make_golden.py full)."""
    import os
    import sys
    here = os.path.join(os.path.dirname(__file__), "golden")
    sys.path.insert(0, here)
    from fullsize import full_size_inputs, full_size_positions
    g = np.load(os.path.join(here, "reference_full_size.npz"))
    mic, ref = full_size_inputs()

    def check(name, arr, tol, sums=()):
        assert list(arr.shape) == list(g[name + "_shape"])
        want = g[name + "_vals"]
        got = arr.ravel()[full_size_positions(arr.shape)]
        assert np.abs(got - want).max() <= tol * max(1.0, np.abs(want).max())
        for key, fn in sums:
            assert np.abs(fn(arr.astype(np.float64)) - g[key]).max() <= 1e-4 * max(1.0, np.abs(g[key]).max())

    tm, tr = _cuda(mic), _cuda(ref)
    s = A.ConvSTFT(512, 256, 512, "hann", "complex")(tm)
    check("stft", s.cpu().numpy(), 2e-5, [("stft_sum_t", lambda a: a.sum(axis=2)), ("stft_pow_c", lambda a: (a ** 2).sum(axis=1))])
    y = A.ConviSTFT(512, 256, 512, "hann", "complex")(s)
    check("istft", y.cpu().numpy(), 2e-5, [("istft_pow", lambda a: (a ** 2).sum(axis=(1, 2)))])
    erb = torch.from_numpy(O.erb_filterbank()).float().cuda()
    f = A.stage2_features(tm, tr, erb)
    check("feat", f.cpu().numpy(), 2e-4, [("feat_sum_t", lambda a: a.sum(axis=1))])
    w = np.load(os.path.join(here, "reference_stage2.npz"))
    net = A.LittleNetInference({k[2:]: w[k] for k in w.files if k.startswith("w_")}, w["erb"])
    o = net(tm, tr)
    check("net", o.cpu().numpy(), 3e-4, [("net_pow", lambda a: (a ** 2).sum(axis=-1).reshape(-1))])


def test_stft_3d_input_and_ctor_errors():
    x = torch.randn(2, 1, 2048, device="cuda")
    s = A.ConvSTFT(512, 256, 512, "hann", "complex")(x)
    assert tuple(s.shape) == (2, 514, 9)
    with pytest.raises(NotImplementedError):
        A.ConvSTFT(400, 100, 512, "hamming", "real")


# ---------------------------------------------------------------------------------------------
# size-independent properties at BASELINE.json's full size (configs[1]: 1024 x 10 s)
# ---------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def full_batch():
    g = torch.Generator(device="cuda").manual_seed(7)
    B, L = 1024, 160000
    far = 0.1 * torch.randn(B, L, device="cuda", generator=g)
    mic = 0.4 * torch.roll(far, 100, dims=1) + 0.2 * torch.roll(far, 300, dims=1)
    mic += 0.001 * torch.randn(B, L, device="cuda", generator=g)
    return far, mic


def test_full_size_zero_far_end_is_identity(full_batch):
    far, mic = full_batch
    err = A.stage1_aec(torch.zeros_like(far), mic)
    assert float((err - mic).abs().max()) < 5e-6


def test_full_size_nlms_linearity_and_determinism(full_batch):
    far, mic = full_batch
    g = torch.Generator(device="cuda").manual_seed(8)
    mic2 = 0.05 * torch.randn(mic.shape, device="cuda", generator=g)
    cfg = A.Stage1Config()
    a = A.stage1_aec(far, mic, cfg)
    b = A.stage1_aec(far, mic2, cfg)
    c = A.stage1_aec(far, 0.5 * mic - 2.0 * mic2, cfg)
    assert float((c - (0.5 * a - 2.0 * b)).abs().max()) < 2e-5
    assert torch.equal(a, A.stage1_aec(far, mic, cfg))                       # run-to-run determinism
    assert torch.isfinite(a).all()


def test_full_size_batch_invariance_and_erle(full_batch):
    far, mic = full_batch
    cfg = A.Stage1Config(erle_skip_hops=125)
    err, erle = A.stage1_aec(far, mic, cfg, return_erle=True)
    sub, sub_erle = A.stage1_aec(far[500:503].clone(), mic[500:503].clone(), cfg, return_erle=True)
    assert torch.equal(err[500:503], sub) and torch.equal(erle[500:503], sub_erle)
    # echo path is a 2-tap filter well inside the 4-partition span: the canceller must converge
    assert float(erle.min()) > 8.0     # STFT-domain filter without cross-band terms: ~10 dB on white noise
    ref = 10 * torch.log10((mic[:, 32000:] ** 2).sum(1) / (err[:, 32000:] ** 2).sum(1))
    assert float((ref - erle).abs().max()) < 0.05


def test_full_size_echo_plus_error_reconstructs_mic(full_batch):
    far, mic = full_batch
    err, echo = A.stage1_aec(far[:256], mic[:256], A.Stage1Config(algo=A.ALGO_KALMAN), return_echo=True)
    assert float((err + echo - mic[:256]).abs().max()) < 2e-5
