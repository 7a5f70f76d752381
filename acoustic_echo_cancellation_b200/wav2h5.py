"""Drop-in for the reference's wav -> h5 generators with the stage-1 canceller inserted.

    create_h5_train(args)  <-> Stage2_lhm/generate_h5files/train_wav2h5.py:10-52
    create_h5_test(args)   <-> Stage2_lhm/generate_h5files/test_wav2h5.py:10-64

Same argparse flags (``--train_path``/``--val_path``, ``--h5_path``, ``--list_path``, ``--sr``), same
file names (``tr/tr_<idx>.ex``, ``tt/test.ex``, ``tr_list.txt``, ``tt_list.txt``, ``filename.txt``) and the
same four dataset keys, so the reference readers (Stage2_lhm/scripts/train1.py:33-40,
scripts/test.py:20-33) keep working; two datasets are ADDED beside them: ``stage1_error`` and
``stage1_echo``.  In ``test.ex`` they go INSIDE each numbered group because ``ValidateDataset`` counts
root members (scripts/test.py:23).

Utterances are batched through the fused CUDA kernel (host-buffer C ABI); with ``torch.distributed``
initialised each rank converts a contiguous shard of the file list and rank 0 writes the merged lists.
``h5py`` is required to write (it is absent from the build image; the import is deferred so the
module can be imported and its batching logic tested without it).  wav decoding uses ``librosa`` when
present (as the reference does), else ``scipy.io.wavfile`` + polyphase resampling.
"""
from __future__ import annotations

import argparse
import glob
import os
from typing import Callable, Dict, List, Optional

import numpy as np

KEYS = ("nearend_speech", "nearend_mic", "farend_speech", "echo")
WAV_PATTERNS = {
    "nearend_speech": "nearend_speech_fileid_{idx}.wav",
    "nearend_mic": "nearend_mic_fileid_{idx}.wav",
    "farend_speech": "farend_speech_fileid_{idx}.wav",
    "echo": "echo_fileid_{idx}.wav",
}


def load_wav(path: str, sr: int) -> np.ndarray:
    """mono float32 at ``sr`` (what ``librosa.load(path, sr=sr)`` returns, train_wav2h5.py:20)."""
    try:
        import librosa  # type: ignore

        y, _ = librosa.load(path, sr=sr)
        return y.astype(np.float32)
    except ImportError:
        pass
    from scipy.io import wavfile
    from scipy.signal import resample_poly

    rate, data = wavfile.read(path)
    if data.dtype.kind == "i":
        data = data.astype(np.float32) / float(2 ** (8 * data.dtype.itemsize - 1))
    elif data.dtype.kind == "u":
        data = (data.astype(np.float32) - 128.0) / 128.0
    data = data.astype(np.float32)
    if data.ndim == 2:
        data = data.mean(axis=1)
    if rate != sr:
        g = np.gcd(int(rate), int(sr))
        data = resample_poly(data, sr // g, rate // g).astype(np.float32)
    return data


def list_utterance_ids(folder: str) -> List[str]:
    """ids in the order ``glob`` yields them (train_wav2h5.py:13-17)."""
    ids = []
    for p in glob.glob(os.path.join(folder, "nearend_speech_fileid_*.wav")):
        ids.append(os.path.basename(p).split(".wav")[0].split("_")[-1])
    return ids


def as_pcm16(x: np.ndarray) -> Optional[np.ndarray]:
    """The int16 samples ``x`` was decoded from, if ``x`` is exactly ``pcm / 32768`` (what ``librosa.load`` /
    ``wavfile.read`` give for a 16-bit PCM wav at its native rate, train_wav2h5.py:20-23), else None.  Lossless by
    construction: the check is ``x * 32768`` being integers inside the int16 range."""
    q = np.asarray(x, dtype=np.float32) * np.float32(32768.0)
    if q.size and (np.any(q != np.rint(q)) or q.max() > 32767.0 or q.min() < -32768.0):
        return None
    return q.astype(np.int16)


def _stage1_batch(far: List[np.ndarray], mic: List[np.ndarray], runner: Callable):
    """Zero-pad a ragged list to [B, Lmax] (like collate_fn, train1.py:52-61), run, un-pad.

    When every far-end and microphone signal of the batch is exact 16-bit PCM (the usual case: the corpus is wav files)
    the batch goes to the runner as int16, which the host-buffer C ABI uploads as such and converts on the GPU
    (``aec_stage1_run_host_pcm16``): half the host-to-device bytes of a PCIe-bound pipeline, bit-identical results."""
    n = np.array([len(x) for x in far], dtype=np.int64)
    lmax = int(n.max())
    pcm = [(as_pcm16(f), as_pcm16(m[:len(f)])) for f, m in zip(far, mic)]
    use_pcm = all(a is not None and b is not None for a, b in pcm)
    dt = np.int16 if use_pcm else np.float32
    fa = np.zeros((len(far), lmax), dtype=dt)
    mi = np.zeros((len(far), lmax), dtype=dt)
    for i, (f, m) in enumerate(zip(far, mic)):
        fa[i, :len(f)] = pcm[i][0] if use_pcm else f
        mm = pcm[i][1] if use_pcm else m[:len(f)]
        mi[i, :len(mm)] = mm
    err, echo = runner(fa, mi, n)
    return [err[i, :n[i]] for i in range(len(far))], [echo[i, :n[i]] for i in range(len(far))]


def default_runner(cfg=None, slice_utterances: int = 256, device: int = 0) -> Callable:
    """Runner backed by ``aec_stage1_run_host`` (no CPU fallback)."""
    from .stage1 import HostPipeline, Stage1Config

    cfg = cfg or Stage1Config()
    state: Dict[str, object] = {}

    def run(far: np.ndarray, mic: np.ndarray, n: np.ndarray):
        pipe = state.get("pipe")
        if pipe is None or pipe.max_samples < far.shape[1]:
            pipe = HostPipeline(slice_utterances, far.shape[1], device)
            state["pipe"] = pipe
        err = np.empty(far.shape, dtype=np.float32)          # (inputs may be int16 PCM; outputs are float32)
        echo = np.empty(far.shape, dtype=np.float32)
        pipe.run(far, mic, cfg, n_samples=n, err=err, echo=echo)
        return err, echo

    return run


def _h5py():
    try:
        import h5py  # type: ignore
    except ImportError as e:  # pragma: no cover - depends on the box
        raise ImportError("writing the reference's .ex (HDF5) files needs h5py, which is not installed") from e
    return h5py


def _shard(ids: List[str]):
    import torch.distributed as dist

    from .sharding import shard_range

    if dist.is_available() and dist.is_initialized():
        lo, hi = shard_range(len(ids), dist.get_rank(), dist.get_world_size())
        return ids[lo:hi], dist.get_rank()
    return ids, 0


def create_h5_train(args, runner: Optional[Callable] = None, batch: int = 256, h5=None):
    """``create_h5(args)`` of train_wav2h5.py with stage 1 inserted.  One ``tr_<idx>.ex`` per utterance."""
    h5 = h5 or _h5py()
    runner = runner or default_runner()
    ids, rank = _shard(list_utterance_ids(args.train_path))
    os.makedirs(os.path.join(args.h5_path, "tr"), exist_ok=True)
    train_list: List[str] = []
    for b0 in range(0, len(ids), batch):
        chunk = ids[b0:b0 + batch]
        sig = {k: [load_wav(os.path.join(args.train_path, WAV_PATTERNS[k].format(idx=i)), args.sr) for i in chunk]
               for k in KEYS}
        errs, echos = _stage1_batch(sig["farend_speech"], sig["nearend_mic"], runner)
        for j, idx in enumerate(chunk):
            name = os.path.join(args.h5_path, "tr", "tr_" + idx + ".ex")
            train_list.append(str(name))
            w = h5.File(name, "w")
            for k in KEYS:                                                   # train_wav2h5.py:39-42
                a = sig[k][j].astype(np.float32)
                w.create_dataset(k, data=a, shape=a.shape, chunks=True)
            w.create_dataset("stage1_error", data=errs[j], shape=errs[j].shape, chunks=True)
            w.create_dataset("stage1_echo", data=echos[j], shape=echos[j].shape, chunks=True)
            w.close()
    from .sharding import merge_filelists

    merged = merge_filelists(train_list)
    if rank == 0:
        with open(os.path.join(args.list_path, "tr_list.txt"), "w") as f:     # train_wav2h5.py:48-51
            f.write("\n".join(merged))
    return merged


def create_h5_test(args, runner: Optional[Callable] = None, batch: int = 256, h5=None, filename: str = "test.ex"):
    """``create_h5(args)`` of test_wav2h5.py with stage 1 inserted.  One file, one numbered group per
    utterance; single writer (rank 0 semantics: call it on one rank)."""
    h5 = h5 or _h5py()
    runner = runner or default_runner()
    ids = list_utterance_ids(args.val_path)
    os.makedirs(os.path.join(args.h5_path, "tt"), exist_ok=True)
    path = os.path.join(args.h5_path, "tt", filename)
    w = h5.File(path, "w")
    count = 0
    for b0 in range(0, len(ids), batch):
        chunk = ids[b0:b0 + batch]
        sig = {k: [load_wav(os.path.join(args.val_path, WAV_PATTERNS[k].format(idx=i)), args.sr) for i in chunk]
               for k in KEYS}
        errs, echos = _stage1_batch(sig["farend_speech"], sig["nearend_mic"], runner)
        for j in range(len(chunk)):
            g = w.create_group(str(count))                                   # test_wav2h5.py:44
            for k in KEYS:                                                   # test_wav2h5.py:45-48
                a = sig[k][j].astype(np.float32)
                g.create_dataset(k, data=a, shape=a.shape, chunks=True)
            g.create_dataset("stage1_error", data=errs[j], shape=errs[j].shape, chunks=True)
            g.create_dataset("stage1_echo", data=echos[j], shape=echos[j].shape, chunks=True)
            count += 1
    w.close()
    with open(os.path.join(args.list_path, "tt_list.txt"), "w") as f:         # test_wav2h5.py:55-58
        f.write(str(path))
    with open(os.path.join(args.list_path, "filename.txt"), "w") as f:        # test_wav2h5.py:60-62
        f.write("\n".join(ids))
    return path


def build_parser(kind: str) -> argparse.ArgumentParser:
    """The reference's flags and defaults (train_wav2h5.py:55-73, test_wav2h5.py:67-86)."""
    p = argparse.ArgumentParser(description=f"wav -> h5 ({kind}) with the stage-1 echo canceller",
                                formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    if kind == "train":
        p.add_argument("--train_path", type=str, default="/data/lihaoming/datasets/synthetic/train_set")
    else:
        p.add_argument("--val_path", type=str, default="/data/lihaoming/datasets/synthetic/test_set")
    p.add_argument("--h5_path", type=str, default="/data/lihaoming/datasets/synthetic/h5")
    p.add_argument("--list_path", type=str, default="../examples/filelists")
    p.add_argument("--sr", type=int, default=16000)
    return p


def main(kind: str = "train", argv=None):
    args = build_parser(kind).parse_args(argv)
    os.makedirs(args.h5_path, exist_ok=True)
    return create_h5_train(args) if kind == "train" else create_h5_test(args)
