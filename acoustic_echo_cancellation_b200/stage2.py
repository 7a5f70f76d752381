"""Stage-2 residual-echo suppressor INFERENCE on the stage-1 output (the caller side of the hot path).

    LittleNetInference  <-> Little_net.forward, inference branch
                            (Stage2_lhm/scripts/network/ERB.py:252-316; used by scripts/test.py:149-169)

Three kernels of ``libaec_b200.so`` replace the module's conv-STFTs, cuDNN GRU, matmuls and transposed-conv
iSTFT: ``aec_features`` (ERB.py:254-290), ``aec_stage2_mask`` (ERB.py:293-304) and ``aec_stage2_synth``
(ERB.py:306-316).  Weights are the reference module's own ``state_dict`` (``gru1.*``, ``linear1.*``,
``linear2.*``); training stays with the reference (no autograd here).
"""
from __future__ import annotations

import ctypes as C
from typing import Mapping

import numpy as np
import torch

from . import _lib
from .spectral import batch_shift, stage2_features
from .stage1 import _require_cuda_f32, _stream_ptr, num_frames, out_samples

_KEYS = {
    "gru_w_ih": ("gru1.weight_ih_l0", (96, 64)), "gru_w_hh": ("gru1.weight_hh_l0", (96, 32)),
    "gru_b_ih": ("gru1.bias_ih_l0", (96,)), "gru_b_hh": ("gru1.bias_hh_l0", (96,)),
    "lin1_w": ("linear1.weight", (32, 64)), "lin1_b": ("linear1.bias", (32,)),
    "lin2_w": ("linear2.weight", (32, 32)), "lin2_b": ("linear2.bias", (32,)),
}


class LittleNetInference:
    """``out_wav = net(mic, ref)`` with ``net = LittleNetInference(state_dict, erb, device)``.

    ``state_dict`` maps the reference's parameter names (dots or underscores) to tensors / arrays;
    ``erb`` is the [257, 32] bank of ``EquivalentRectangularBandwidth(...).filters``."""

    def __init__(self, state_dict: Mapping[str, object], erb, device="cuda"):
        self.device = torch.device(device)
        self._w = {}
        for field, (name, shape) in _KEYS.items():
            v = state_dict.get(name, state_dict.get(name.replace(".", "_")))
            if v is None:
                raise KeyError(f"state_dict lacks {name}")
            t = torch.as_tensor(np.asarray(v) if not isinstance(v, torch.Tensor) else v).to(
                device=self.device, dtype=torch.float32).contiguous()
            if tuple(t.shape) != shape:
                raise ValueError(f"{name} must have shape {shape} (Little_net with 32 ERB bands)")
            self._w[field] = t
        self.erb = torch.as_tensor(np.asarray(erb) if not isinstance(erb, torch.Tensor) else erb).to(
            device=self.device, dtype=torch.float32).contiguous()
        if tuple(self.erb.shape) != (257, 32):
            raise ValueError("erb must be [257, 32]")
        self._cw = _lib.Stage2Weights(**{k: v.data_ptr() for k, v in self._w.items()})

    def __call__(self, mic: torch.Tensor, ref: torch.Tensor, in_norm: bool = True) -> torch.Tensor:
        _require_cuda_f32("mic", mic)
        _require_cuda_f32("ref", ref)
        if mic.shape != ref.shape or mic.dim() != 2:
            raise ValueError("mic and ref must be [B, L]")
        mic, ref = mic.contiguous(), ref.contiguous()
        B, L = mic.shape
        lib = _lib.load()
        # ERB.py:254: the batch-global scalar mean/std is subtracted (torch.std is unbiased); reduced on the device
        # once per signal and passed by device pointer -- the whole forward is asynchronous, no host round trip
        shift_mic = batch_shift(mic) if in_norm else None
        shift_ref = batch_shift(ref) if in_norm else None
        feat = stage2_features(mic, ref, self.erb, in_norm=in_norm,
                               shifts=(shift_mic, shift_ref) if in_norm else None)   # [B, T, 64]
        T = num_frames(L)
        with torch.cuda.device(mic.device):
            est = torch.empty((B, T, 32), dtype=torch.float32, device=mic.device)
            out = torch.empty((B, out_samples(L)), dtype=torch.float32, device=mic.device)
            s = _stream_ptr(mic)
            _lib.check(lib.aec_stage2_mask(feat.data_ptr(), C.byref(self._cw), est.data_ptr(), B, T, 32, s),
                       "aec_stage2_mask")
            if out.numel():
                _lib.check(lib.aec_stage2_synth_dev(mic.data_ptr(), est.data_ptr(), self.erb.data_ptr(), out.data_ptr(),
                                                    B, L, L, out.shape[1], 512, 32,
                                                    shift_mic.data_ptr() if shift_mic is not None else None, s),
                           "aec_stage2_synth_dev")
        return out
