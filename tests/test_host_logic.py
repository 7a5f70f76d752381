"""CPU: host-side logic -- utterance sharding, the world_size-2 metric gather (gloo), ragged
batching and the h5 schema of the create_h5 drop-in (with an in-memory stand-in for h5py, which is
not installed in the build image)."""
import os
import types

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from acoustic_echo_cancellation_b200 import hostutil, sharding, wav2h5


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 1024, 100000):
        for world in (1, 2, 3, 8):
            spans = [sharding.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.shard_range(4, 2, 2)


def _worker(rank, world, port, n_items, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = sharding.shard_range(n_items, rank, world)
    local = torch.arange(lo, hi, dtype=torch.float32) * 0.5      # stands for per-utterance ERLE
    full = sharding.gather_metrics(local, n_items)
    full_async, work = sharding.gather_metrics(local, n_items, async_op=True)    # overlapped form used by bench.py
    if work is not None:
        work.wait()
    assert torch.equal(full, full_async)
    even, work = sharding.gather_metrics(torch.full((3,), float(rank)), 3 * world, async_op=True)   # equal shards
    if work is not None:
        work.wait()
    assert even.tolist() == [float(r) for r in range(world) for _ in range(3)]
    files = sharding.merge_filelists([f"tr_{i}.ex" for i in range(lo, hi)])
    ret[rank] = (full.tolist(), files)
    dist.destroy_process_group()


def test_gather_metrics_world_size_2_gloo():
    n_items, world = 7, 2
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(world, port, n_items, ret), nprocs=world, join=True)
    for r in range(world):
        full, files = ret[r]
        assert full == [0.5 * i for i in range(n_items)]
        assert files == [f"tr_{i}.ex" for i in range(n_items)]


def test_gather_metrics_single_process_passthrough():
    x = torch.arange(5, dtype=torch.float32)
    assert torch.equal(sharding.gather_metrics(x, 5), x)
    y, work = sharding.gather_metrics(x, 5, async_op=True)
    assert work is None and torch.equal(y, x)


def test_ragged_batching_pads_and_unpads():
    far = [np.arange(5, dtype=np.float32), np.arange(3, dtype=np.float32)]
    mic = [np.ones(5, dtype=np.float32), np.ones(3, dtype=np.float32)]
    seen = {}

    def runner(fa, mi, n):
        seen["shape"], seen["n"] = fa.shape, n.tolist()
        assert (fa[1, 3:] == 0).all()
        return fa + 1, mi * 2

    errs, echos = wav2h5._stage1_batch(far, mic, runner)
    assert seen == {"shape": (2, 5), "n": [5, 3]}
    assert [len(e) for e in errs] == [5, 3] and np.array_equal(errs[1], far[1] + 1)
    assert np.array_equal(echos[0], 2 * mic[0])


class _FakeGroup(dict):
    def create_dataset(self, name, data=None, shape=None, chunks=None):
        assert chunks is True and shape == data.shape and data.dtype == np.float32
        self[name] = np.array(data)

    def create_group(self, name):
        g = _FakeGroup()
        self[name] = g
        return g


class _FakeH5:
    files = {}

    def File(self, name, mode):
        assert mode == "w"
        f = _FakeGroup()
        f.close = lambda: None
        self.files[name] = f
        return f


def _write_wavs(folder, ids, n=2000, sr=16000):
    from scipy.io import wavfile

    rng = np.random.default_rng(0)
    for i in ids:
        for k, pat in wav2h5.WAV_PATTERNS.items():
            x = (rng.standard_normal(n + 10 * int(i)) * 3000).astype(np.int16)
            wavfile.write(os.path.join(folder, pat.format(idx=i)), sr, x)


def test_create_h5_train_and_test_schema(tmp_path):
    wav_dir, h5_dir, list_dir = tmp_path / "wav", tmp_path / "h5", tmp_path / "lists"
    for d in (wav_dir, h5_dir, list_dir):
        d.mkdir()
    _write_wavs(str(wav_dir), ["3", "11"])
    def runner(fa, mi, n):            # stand-in for the CUDA runner: like the C ABI it takes float32 or 16-bit PCM
        if fa.dtype == np.int16:      # (the test wavs are 16-bit, so this is the branch create_h5 takes)
            fa, mi = fa.astype(np.float32) / np.float32(32768), mi.astype(np.float32) / np.float32(32768)
        return fa * np.float32(0.5), mi * np.float32(0.25)

    fake = _FakeH5()
    args = types.SimpleNamespace(train_path=str(wav_dir), h5_path=str(h5_dir), list_path=str(list_dir), sr=16000)
    paths = wav2h5.create_h5_train(args, runner=runner, batch=1, h5=fake)
    assert sorted(os.path.basename(p) for p in paths) == ["tr_11.ex", "tr_3.ex"]
    assert open(list_dir / "tr_list.txt").read().split("\n") == paths
    for p in paths:
        f = fake.files[p]
        assert set(f) == {"nearend_speech", "nearend_mic", "farend_speech", "echo", "stage1_error", "stage1_echo"}
        assert f["stage1_error"].shape == f["farend_speech"].shape
        assert np.allclose(f["stage1_error"], 0.5 * f["farend_speech"])
    args = types.SimpleNamespace(val_path=str(wav_dir), h5_path=str(h5_dir), list_path=str(list_dir), sr=16000)
    path = wav2h5.create_h5_test(args, runner=runner, batch=2, h5=fake)
    root = fake.files[path]
    assert sorted(root) == ["0", "1"]                     # ONLY numbered groups at the root (test.py:23)
    assert "stage1_error" in root["0"] and "nearend_mic" in root["1"]
    assert open(list_dir / "tt_list.txt").read() == path
    assert len(open(list_dir / "filename.txt").read().split("\n")) == 2


def test_cli_flags_match_the_reference():
    a = wav2h5.build_parser("train").parse_args([])
    assert (a.train_path, a.h5_path, a.list_path, a.sr) == (
        "/data/lihaoming/datasets/synthetic/train_set", "/data/lihaoming/datasets/synthetic/h5",
        "../examples/filelists", 16000)
    b = wav2h5.build_parser("test").parse_args(["--val_path", "x", "--sr", "8000"])
    assert b.val_path == "x" and b.sr == 8000


def test_cpulist_parser_and_bind_is_harmless_without_gpu():
    assert hostutil._parse_cpulist("0-3,8,10-11\n") == [0, 1, 2, 3, 8, 10, 11]
    assert hostutil._parse_cpulist("") == []
    if not torch.cuda.is_available():
        assert hostutil.gpu_local_cpus(0) is None and hostutil.bind_to_gpu_numa(0) is None


def test_stage1_batch_sends_exact_pcm16_as_int16_and_everything_else_as_float32():
    rng = np.random.default_rng(0)
    pcm = [rng.integers(-32768, 32768, size=n, dtype=np.int16) for n in (700, 512)]
    far = [p.astype(np.float32) / 32768.0 for p in pcm]            # what librosa.load returns for a 16-bit wav
    mic = [np.roll(f, 3) for f in far]
    seen = {}

    def runner(fa, mi, n):
        seen["dtype"], seen["far"], seen["n"] = fa.dtype, fa.copy(), n.copy()
        z = np.zeros(fa.shape, dtype=np.float32)
        return z, z

    errs, _ = wav2h5._stage1_batch(far, mic, runner)
    assert seen["dtype"] == np.int16 and [len(e) for e in errs] == [700, 512]
    assert np.array_equal(seen["far"][0, :700], pcm[0]) and np.array_equal(seen["far"][1, :512], pcm[1])
    assert (seen["far"][1, 512:] == 0).all()
    far[1] = far[1] + np.float32(1e-6)                             # not representable as PCM16 any more
    wav2h5._stage1_batch(far, mic, runner)
    assert seen["dtype"] == np.float32
    assert wav2h5.as_pcm16(np.array([1.0], dtype=np.float32)) is None          # +1.0 would be 32768: out of range
    assert wav2h5.as_pcm16(np.array([-1.0, 0.5], dtype=np.float32)).tolist() == [-32768, 16384]
