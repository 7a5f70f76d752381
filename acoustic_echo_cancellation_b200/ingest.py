"""Batched wav ingest for the h5 generators.

The reference decodes four wav files per utterance with ``librosa.load`` in a serial loop
(Stage2_lhm/generate_h5files/train_wav2h5.py:13-23, test_wav2h5.py:21-32, val_wav2h5.py:24-36) -- on the
real pipeline that loop, not the filter, is where the time goes.  Here a batch of utterances is decoded by a
thread pool straight into page-locked batch buffers:

* fast path -- 16-bit PCM, mono, already at ``sr`` (what the synthetic AEC corpora are): the RIFF header is
  parsed here, the sample bytes are ``readinto`` the int16 row of the batch buffer (no float round trip, no
  intermediate array; the read releases the GIL), and the batch goes to the GPU as int16
  (``aec_stage1_run_host_pcm16``), where ``x / 32768`` -- exactly what ``librosa.load`` returns for such a
  file -- is applied.
* anything else (other sample formats, stereo, another rate) falls back to ``load_wav`` (librosa when present,
  else scipy + polyphase resampling) and the batch travels as float32.

Host-side plumbing only; no arithmetic on the samples happens here.
"""
from __future__ import annotations

import os
import struct
from concurrent.futures import ThreadPoolExecutor
from dataclasses import dataclass
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np


@dataclass
class WavInfo:
    rate: int
    channels: int
    bits: int
    fmt: int            # 1 = integer PCM, 3 = IEEE float, 0xFFFE = extensible (sub-format in `fmt_sub`)
    frames: int
    data_offset: int

    def fast(self, sr: int) -> bool:
        return self.fmt == 1 and self.bits == 16 and self.channels == 1 and self.rate == sr


def probe_wav(path: str) -> WavInfo:
    """Parse the RIFF/WAVE chunk list up to the ``data`` chunk (a few dozen bytes; no sample is read)."""
    with open(path, "rb") as f:
        head = f.read(12)
        if len(head) < 12 or head[:4] != b"RIFF" or head[8:12] != b"WAVE":
            raise ValueError(f"{path}: not a RIFF/WAVE file")
        fmt = None
        while True:
            ch = f.read(8)
            if len(ch) < 8:
                raise ValueError(f"{path}: no data chunk")
            cid, size = ch[:4], struct.unpack("<I", ch[4:])[0]
            if cid == b"fmt ":
                body = f.read(size + (size & 1))
                tag, nch, rate, _, _, bits = struct.unpack("<HHIIHH", body[:16])
                if tag == 0xFFFE and size >= 26:
                    tag = struct.unpack("<H", body[24:26])[0]
                fmt = (tag, nch, rate, bits)
            elif cid == b"data":
                if fmt is None:
                    raise ValueError(f"{path}: data chunk before fmt chunk")
                tag, nch, rate, bits = fmt
                off = f.tell()
                avail = os.fstat(f.fileno()).st_size - off
                size = min(size, avail)           # (streamed files carry 0xFFFFFFFF here)
                return WavInfo(rate, nch, bits, tag, size // max(nch * bits // 8, 1), off)
            else:
                f.seek(size + (size & 1), os.SEEK_CUR)


def read_pcm16_into(path: str, info: WavInfo, row: np.ndarray) -> int:
    """Read the samples of a fast-path file into ``row`` (int16, contiguous, >= info.frames long); zero the rest."""
    n = info.frames
    with open(path, "rb", buffering=0) as f:
        f.seek(info.data_offset)
        mv = memoryview(row[:n]).cast("B")
        got = 0
        while got < 2 * n:
            k = f.readinto(mv[got:])
            if not k:
                raise IOError(f"{path}: truncated data chunk")
            got += k
    row[n:] = 0
    return n


def _native():
    """libaec_b200.so's batched reader (``aec_wav_probe_batch`` / ``aec_wav_read_pcm16_batch``), or None when the
    library has not been built (host-only unit tests of this module then use the Python reader above)."""
    try:
        from . import _lib

        return _lib, _lib.load()
    except (ImportError, OSError, AttributeError):
        return None


def _c_paths(paths: Sequence[str]):
    import ctypes as C

    arr = (C.c_char_p * len(paths))()
    arr[:] = [os.fsencode(p) for p in paths]
    return arr


def probe_batch(paths: Sequence[str], threads: int) -> Optional[List[WavInfo]]:
    """Headers of many files at once through the native reader; None if it is unavailable."""
    nat = _native()
    if nat is None or not paths:
        return None if nat is None else []
    _lib, lib = nat
    infos = (_lib.WavInfo * len(paths))()
    rc = lib.aec_wav_probe_batch(_c_paths(paths), len(paths), infos, int(threads))
    if rc == -6:
        raise IOError("aec_wav_probe_batch: a wav file could not be opened or read")
    if rc != 0:
        raise ValueError("aec_wav_probe_batch: not a RIFF/WAVE file")
    return [WavInfo(i.rate, i.channels, i.bits, i.format, i.frames, i.data_offset) for i in infos]


def read_pcm16_batch(paths: Sequence[str], dst: np.ndarray, sr: int, threads: int) -> Optional[np.ndarray]:
    """Native multi-threaded read of fast-path files into the rows of ``dst`` ([n, cols] int16, contiguous rows).
    Returns the true frame counts, or None when a file is not 16-bit mono PCM at ``sr`` (caller falls back)."""
    _lib, lib = _native()
    n = len(paths)
    frames = np.zeros(n, dtype=np.int64)
    assert dst.dtype == np.int16 and dst.strides[1] == 2 and dst.shape[0] >= n
    rc = lib.aec_wav_read_pcm16_batch(_c_paths(paths), n, dst.ctypes.data, dst.strides[0] // 2, dst.shape[1],
                                      frames.ctypes.data, int(sr), int(threads))
    if rc == -2:
        return None
    if rc != 0:
        raise IOError(f"aec_wav_read_pcm16_batch failed ({rc})")
    return frames


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


def pcm16_to_float32(x: np.ndarray) -> np.ndarray:
    """``x / 32768`` in float32: the value ``librosa.load`` gives for a 16-bit sample."""
    return x.astype(np.float32) * np.float32(1.0 / 32768.0)


@dataclass
class DecodedBatch:
    """One batch of utterances as the generators need it.

    ``far`` / ``mic`` are [nb, lmax] batch buffers (int16 when ``pcm16`` else float32; page-locked when the
    allocator is), zero-padded; ``n`` the true lengths (of the far-end file, the filter's clock);
    ``signals[key][j]`` the per-utterance arrays to be STORED (float32 or int16 views, un-padded)."""

    far: np.ndarray
    mic: np.ndarray
    n: np.ndarray
    pcm16: bool
    signals: Dict[str, List[np.ndarray]]


class BatchDecoder:
    """Decodes batches of (far, mic, + stored-only signals) with a thread pool into reusable batch buffers.

    ``alloc(shape, dtype)`` supplies the batch buffers (``stage1.pinned_empty`` on the GPU path; ``np.empty``
    in host-only tests).  ``sets`` buffer sets are cycled, so that a batch can still be on the GPU / being
    written while the next ones are decoded."""

    def __init__(self, sr: int, threads: int = 8, alloc: Optional[Callable] = None, sets: int = 3,
                 native: Optional[bool] = None):
        self.sr = int(sr)
        self.threads = max(1, int(threads))
        self.native = (_native() is not None) if native is None else bool(native)
        self.pool = ThreadPoolExecutor(max(1, int(threads)), thread_name_prefix="aec-wav")
        self.alloc = alloc or (lambda shape, dtype: np.empty(shape, dtype=dtype))
        self.sets = [dict() for _ in range(max(1, sets))]
        self._next = 0

    def close(self):
        self.pool.shutdown(wait=True)
        self.sets = []

    def _buffer(self, slot: dict, name: str, nb: int, lmax: int, dtype, upload: bool = True) -> np.ndarray:
        """[nb, lmax] view of a reusable batch buffer of this set; ``upload=False``: stored-only signals, plain
        (pageable) memory is enough."""
        key = (name, np.dtype(dtype).str)
        buf = slot.get(key)
        if buf is None or buf.shape[0] < nb or buf.shape[1] < lmax:
            rows = max(nb, buf.shape[0] if buf is not None else 0)
            cols = max(lmax, buf.shape[1] if buf is not None else 0)
            buf = self.alloc((rows, cols), dtype) if upload else np.empty((rows, cols), dtype=dtype)
            slot[key] = buf
        # [nb, lmax] view with the buffer's row pitch would not be contiguous; the host entry takes a stride
        return buf[:nb, :lmax]

    def _decode_native(self, slot, far_paths, mic_paths, extra) -> Optional[DecodedBatch]:
        """All-native form of the fast path: one probe call, one read call per signal (C++ threads, no GIL)."""
        nb = len(far_paths)
        infos = probe_batch(list(far_paths) + list(mic_paths), self.threads)
        if not all(a.fast(self.sr) for a in infos):
            return None
        fi, mi = infos[:nb], infos[nb:]
        n = np.array([a.frames for a in fi], dtype=np.int64)
        lmax = int(max(int(n.max()), 1))
        far = self._buffer(slot, "far", nb, lmax, np.int16)
        mic = self._buffer(slot, "mic", nb, lmax, np.int16)
        if read_pcm16_batch(far_paths, far, self.sr, self.threads) is None:
            return None
        if read_pcm16_batch(mic_paths, mic, self.sr, self.threads) is None:
            return None
        signals: Dict[str, List[np.ndarray]] = {"__far__": [far[j, :n[j]] for j in range(nb)], "__mic__": []}
        for j in range(nb):
            if mi[j].frames > n[j]:
                mic[j, n[j]:] = 0                       # the uploaded row follows the far-end clock
            signals["__mic__"].append(mic[j, :n[j]] if mi[j].frames == n[j] else None)
        ragged = [j for j in range(nb) if signals["__mic__"][j] is None]
        for key, paths in list(extra.items()) + ([("__mic__", [mic_paths[j] for j in ragged])] if ragged else []):
            ki = probe_batch(paths, self.threads)
            if not all(a.fast(self.sr) for a in ki):
                out = list(self.pool.map(lambda p: load_wav(p, self.sr), paths))
            else:
                cols = int(max([a.frames for a in ki] + [1]))
                buf = self._buffer(slot, "store_" + key, len(paths), cols, np.int16, upload=False)
                fr = read_pcm16_batch(paths, buf, self.sr, self.threads)
                out = [buf[j, :fr[j]] for j in range(len(paths))]
            if key == "__mic__":
                for j, a in zip(ragged, out):
                    signals["__mic__"][j] = a
            else:
                signals[key] = out
        return DecodedBatch(far=far, mic=mic, n=n, pcm16=True, signals=signals)

    def decode(self, far_paths: Sequence[str], mic_paths: Sequence[str],
               extra: Dict[str, Sequence[str]]) -> DecodedBatch:
        """``extra`` maps a storage key (e.g. ``nearend_speech``) to its paths; those are decoded but only
        stored, never uploaded."""
        nb = len(far_paths)
        slot = self.sets[self._next]
        self._next = (self._next + 1) % len(self.sets)
        if self.native:
            b = self._decode_native(slot, far_paths, mic_paths, extra)
            if b is not None:
                return b
        infos = list(self.pool.map(probe_wav, list(far_paths) + list(mic_paths)))
        fi, mi = infos[:nb], infos[nb:]
        fast = all(a.fast(self.sr) for a in infos)
        signals: Dict[str, List[np.ndarray]] = {}
        if fast:
            n = np.array([a.frames for a in fi], dtype=np.int64)
            lmax = int(max(int(n.max()), 1))
            far = self._buffer(slot, "far", nb, lmax, np.int16)
            mic = self._buffer(slot, "mic", nb, lmax, np.int16)

            def job(j):
                read_pcm16_into(far_paths[j], fi[j], far[j])
                # the microphone row follows the far-end clock: cut or zero-pad to n[j] (collate_fn semantics)
                m = mi[j]
                if m.frames >= lmax:
                    read_pcm16_into(mic_paths[j], WavInfo(m.rate, 1, 16, 1, lmax, m.data_offset), mic[j])
                else:
                    read_pcm16_into(mic_paths[j], m, mic[j])
                if m.frames > n[j]:
                    mic[j, n[j]:] = 0
                return None

            list(self.pool.map(job, range(nb)))
            signals["__far__"] = [far[j, :n[j]] for j in range(nb)]
            # the STORED microphone signal keeps its own length, like the reference's dataset does: a view of the
            # batch row when it equals the far-end length (the corpus case), else re-read in full below
            signals["__mic__"] = [mic[j, :n[j]] if mi[j].frames == n[j] else None for j in range(nb)]
        else:
            fa = list(self.pool.map(lambda p: load_wav(p, self.sr), far_paths))
            ma = list(self.pool.map(lambda p: load_wav(p, self.sr), mic_paths))
            n = np.array([len(x) for x in fa], dtype=np.int64)
            lmax = int(max(int(n.max()), 1))
            far = self._buffer(slot, "far", nb, lmax, np.float32)
            mic = self._buffer(slot, "mic", nb, lmax, np.float32)
            for j in range(nb):
                far[j, :n[j]] = fa[j]
                far[j, n[j]:] = 0
                k = min(len(ma[j]), int(n[j]))
                mic[j, :k] = ma[j][:k]
                mic[j, k:] = 0
            signals["__far__"], signals["__mic__"] = fa, ma

        def stored(path):
            info = probe_wav(path)
            if info.fast(self.sr):
                a = np.empty(info.frames, dtype=np.int16)
                read_pcm16_into(path, info, a)
                return a
            return load_wav(path, self.sr)

        for j in range(nb):
            if signals["__mic__"][j] is None:
                signals["__mic__"][j] = stored(mic_paths[j])
        for key, paths in extra.items():
            signals[key] = list(self.pool.map(stored, paths))
        return DecodedBatch(far=far, mic=mic, n=n, pcm16=fast, signals=signals)


def as_float32(a: np.ndarray) -> np.ndarray:
    """Stored form of a decoded signal: float32, ``/ 32768`` for the int16 fast path."""
    return pcm16_to_float32(a) if a.dtype == np.int16 else np.asarray(a, dtype=np.float32)
