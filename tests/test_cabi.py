"""CPU: the C-ABI shared library loads and exports every symbol include/aec_b200.h declares; the
device-free helpers agree with the reference-pinned golden frame counts; argument errors are
reported as negative codes and a missing GPU is reported loudly (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import acoustic_echo_cancellation_b200 as A
from acoustic_echo_cancellation_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
    src = open(os.path.join(ROOT, "include", "aec_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(aec_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    names = _declared_functions()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/aec_b200.h but not exported"
    assert set(names) == set(_lib.SIGNATURES), set(names) ^ set(_lib.SIGNATURES)


def test_version_and_strerror():
    lib = _lib.load()
    assert lib.aec_version() == 100
    assert lib.aec_strerror(0) == b"ok"
    assert b"invalid" in lib.aec_strerror(-1)


def test_cfg_default_and_struct_layout():
    cfg = _lib.default_cfg(512)
    assert (cfg.frame, cfg.partitions, cfg.algo) == (512, 4, 0)
    assert cfg.mu == pytest.approx(0.5) and cfg.delta == pytest.approx(512e-6)
    assert cfg.kalman_a == pytest.approx(0.999) and cfg.kalman_lambda == pytest.approx(0.9)
    assert C.sizeof(_lib.AecCfg) == 64
    assert _lib.load().aec_cfg_default(C.byref(cfg), 500) == -1


def test_frame_helpers_match_reference_counts(golden):
    for n, t in zip(golden["frame_count_L"], golden["frame_count_T"]):
        assert A.num_frames(int(n)) == int(t)
        assert A.out_samples(int(n)) == (int(t) - 1) * 256
    assert A.num_frames(480000, 1024) == 938


def test_argument_errors_do_not_need_a_gpu():
    lib = _lib.load()
    cfg = _lib.default_cfg(512)
    f = np.zeros(8, dtype=np.float32).ctypes.data
    run = lib.aec_stage1_run
    assert run(f, f, f, None, None, None, 0, 0, 0, 0, C.byref(cfg), None) == 0          # empty batch
    assert run(f, f, f, None, None, None, -1, 8, 8, 8, C.byref(cfg), None) == -1
    assert run(f, f, f, None, None, None, 1, 8, 4, 8, C.byref(cfg), None) == -1          # stride < L
    assert run(None, f, f, None, None, None, 1, 8, 8, 8, C.byref(cfg), None) == -1
    bad = _lib.default_cfg(512, partitions=0)
    assert run(f, f, f, None, None, None, 1, 8, 8, 8, C.byref(bad), None) == -1
    bad = _lib.default_cfg(512, algo=7)
    assert run(f, f, f, None, None, None, 1, 8, 8, 8, C.byref(bad), None) == -1
    assert lib.aec_stft(f, f, 1, 8, 8, 300, None) == -1
    assert lib.aec_features(f, f, f, f, 1, 8, 8, 1024, 32, 0.0, 0.0, None) == -2   # 257-bin ERB config only
    assert lib.aec_istft(f, f, 1, 5, 8, 512, None) == -1                                # out_stride too small


def test_no_cpu_fallback():
    import torch

    with pytest.raises(RuntimeError):
        A.stage1_aec(torch.zeros(1, 1024), torch.zeros(1, 1024))                         # CPU tensors rejected
    if not torch.cuda.is_available():
        lib = _lib.load()
        cfg = _lib.default_cfg(512)
        f = np.zeros(1024, dtype=np.float32).ctypes.data
        rc = lib.aec_stage1_run(f, f, f, None, None, None, 1, 1024, 1024, 1024, C.byref(cfg), None)
        assert rc in (-3, -4)                                                           # loud, not silent
        assert lib.aec_last_cuda_error() != b""


def test_product_package_never_imports_the_oracle():
    """comments may cite the oracle; code may not import, include, link or dlopen it"""
    pkg = os.path.join(ROOT, "acoustic_echo_cancellation_b200")
    py_bad = re.compile(r"^\s*(from|import)\s+\.*oracle|c_oracle|libaec_oracle", re.M)
    c_bad = re.compile(r"#\s*include[^\n]*oracle|libaec_oracle|aec_oracle_stage1", re.M)
    for dirpath, _, files in os.walk(pkg):
        if os.sep + "build" in dirpath:
            continue
        for fn in files:
            path = os.path.join(dirpath, fn)
            if fn.endswith(".py"):
                assert not py_bad.search(open(path).read()), fn
            elif fn.endswith((".cu", ".cuh", ".h")) or fn == "Makefile":
                assert not c_bad.search(open(path).read()), fn
