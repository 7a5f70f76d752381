"""ctypes loader for the C oracle (oracle/csrc/aec_oracle.c).  TEST INFRASTRUCTURE ONLY: imported
by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs, never by the
product package."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libaec_oracle.so")


class OracleCfg(C.Structure):
    _fields_ = [("frame", C.c_int32), ("partitions", C.c_int32), ("algo", C.c_int32),
                ("mu", C.c_float), ("delta", C.c_float), ("kalman_a", C.c_float),
                ("kalman_lambda", C.c_float), ("kalman_c0", C.c_float), ("kalman_eps", C.c_float),
                ("erle_skip_hops", C.c_int32), ("pb_lambda", C.c_float)]


_lib = None


def load(build_if_missing: bool = True) -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH) and build_if_missing:
            subprocess.check_call(["make", "-C", os.path.join(_HERE, "csrc")])
        lib = C.CDLL(LIB_PATH)
        lib.aec_oracle_stage1_f32.restype = C.c_int
        lib.aec_oracle_stage1_f32.argtypes = [C.c_void_p] * 6 + [C.c_int64] * 4 + [C.POINTER(OracleCfg), C.c_int]
        lib.aec_oracle_max_threads.restype = C.c_int
        _lib = lib
    return _lib


def make_cfg(cfg, erle_skip_hops: int = 0) -> OracleCfg:
    """cfg: oracle.aec_oracle.AecConfig"""
    return OracleCfg(cfg.frame, cfg.partitions, cfg.algo, cfg.mu, cfg.delta, cfg.kalman_a, cfg.kalman_lambda,
                     cfg.kalman_c0, cfg.kalman_eps, erle_skip_hops, getattr(cfg, "pb_lambda", 0.5))


def stage1(far: np.ndarray, mic: np.ndarray, cfg, n_samples=None, want_echo: bool = True,
           erle_skip_hops: int = 0, n_threads: int = 0, out=None):
    """float32 C oracle on [B, L] arrays.  Returns dict(err, echo, erle_db, threads).
    ``out`` = (err, echo_or_None, erle) reuses caller buffers (timing runs)."""
    lib = load()
    far = np.ascontiguousarray(far, dtype=np.float32)
    mic = np.ascontiguousarray(mic, dtype=np.float32)
    B, L = far.shape
    if out is not None:
        err, echo, erle = out
    else:
        err = np.empty((B, L), dtype=np.float32)
        echo = np.empty((B, L), dtype=np.float32) if want_echo else None
        erle = np.empty(B, dtype=np.float32)
    ns = None if n_samples is None else np.ascontiguousarray(n_samples, dtype=np.int64)
    c = make_cfg(cfg, erle_skip_hops)
    rc = lib.aec_oracle_stage1_f32(far.ctypes.data, mic.ctypes.data, err.ctypes.data,
                                   echo.ctypes.data if echo is not None else None, erle.ctypes.data,
                                   ns.ctypes.data if ns is not None else None, B, L, L, L, C.byref(c), n_threads)
    if rc < 0:
        raise RuntimeError(f"aec_oracle_stage1_f32 failed: {rc}")
    return {"err": err, "echo": echo, "erle_db": erle, "threads": rc}
