"""Device timing of the stand-alone operators against the reference's formulation run on the same GPU.

The reference computes its STFT as a dense strided convolution with a [2K, 1, N] kernel
(Stage2_lhm/scripts/network/attention_ccrn.py:8-25, 45-52) and its iSTFT as a transposed convolution
(:82-101).  /root/reference is not on the GPU box, so the conv formulation is RESTATED here with torch
(for timing only; parity is pinned separately by the golden vectors)."""
import json
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import acoustic_echo_cancellation_b200 as A  # noqa: E402


def conv_kernels(n=512):
    w = torch.hann_window(n, periodic=True, dtype=torch.float64)
    basis = torch.fft.rfft(torch.eye(n, dtype=torch.float64))           # [n, K]
    k = torch.cat([basis.real, basis.imag], 1).T                        # [2K, n]
    inv = torch.linalg.pinv(k).T
    return (k * w)[:, None, :].float().cuda(), (inv * w)[:, None, :].float().cuda(), w.float().cuda()


def timeit(fn, n=10):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n


def main():
    B, L, n, hop = 128, 160000, 512, 256
    x = 0.1 * torch.randn(B, L, device="cuda")
    y = 0.1 * torch.randn(B, L, device="cuda")
    kf, ki, w = conv_kernels(n)
    stft, istft = A.ConvSTFT(n, hop, n, "hann", "complex"), A.ConviSTFT(n, hop, n, "hann", "complex")
    erb = torch.from_numpy(A.erb_filterbank()).float().cuda()

    def ref_stft(v):
        return F.conv1d(F.pad(v[:, None, :], [n - hop, n - hop]), kf, stride=hop)

    def ref_istft(s):
        o = F.conv_transpose1d(s, ki, stride=hop)
        t = (w[None, :, None] ** 2).repeat(1, 1, s.size(-1))
        coff = F.conv_transpose1d(t, torch.eye(n, device="cuda")[:, None, :], stride=hop)
        return (o / (coff + 1e-8))[..., n - hop:-(n - hop)]

    def ref_feat(m, r):
        out = []
        for v in (m, r):
            s = ref_stft(v)
            mag = torch.sqrt(s[:, :257] ** 2 + s[:, 257:] ** 2 + 1e-9).transpose(1, 2)
            out.append(mag @ erb)
        return torch.cat([out[0], (out[0] - out[1]).abs()], 2)

    s_ours = stft(x)
    res = {
        "shape": [B, L],
        "stft_ms": {"ours": timeit(lambda: stft(x)), "conv_formulation": timeit(lambda: ref_stft(x))},
        "istft_ms": {"ours": timeit(lambda: istft(s_ours)), "conv_formulation": timeit(lambda: ref_istft(s_ours))},
        "features_ms": {"ours": timeit(lambda: A.stage2_features(x, y, erb, in_norm=False)),
                        "conv_formulation": timeit(lambda: ref_feat(x, y))},
        "max_abs_diff": {"stft": float((s_ours - ref_stft(x)).abs().max()),
                         "istft": float((istft(s_ours) - ref_istft(s_ours)).abs().max()),
                         "features": float((A.stage2_features(x, y, erb, in_norm=False) - ref_feat(x, y)).abs().max())},
    }
    # ---- Stage-2 inference (Little_net.forward, ERB.py:252-316) restated with torch modules (cuDNN GRU) ----
    torch.manual_seed(0)
    gru = torch.nn.GRU(64, 32, batch_first=True).cuda().eval()
    lin1, lin2 = torch.nn.Linear(64, 32).cuda().eval(), torch.nn.Linear(32, 32).cuda().eval()
    sd = {"gru1.weight_ih_l0": gru.weight_ih_l0, "gru1.weight_hh_l0": gru.weight_hh_l0, "gru1.bias_ih_l0": gru.bias_ih_l0,
          "gru1.bias_hh_l0": gru.bias_hh_l0, "linear1.weight": lin1.weight, "linear1.bias": lin1.bias,
          "linear2.weight": lin2.weight, "linear2.bias": lin2.bias}
    net = A.LittleNetInference({k: v.detach() for k, v in sd.items()}, erb)

    @torch.no_grad()
    def ref_stage2(m, r):
        m = m - m.mean() / m.std()
        r = r - r.mean() / r.std()
        sm, sr = ref_stft(m), ref_stft(r)
        mm = torch.sqrt(sm[:, :257] ** 2 + sm[:, 257:] ** 2 + 1e-9).transpose(1, 2) @ erb
        mr = torch.sqrt(sr[:, :257] ** 2 + sr[:, 257:] ** 2 + 1e-9).transpose(1, 2) @ erb
        feat = torch.cat([mm, (mm - mr).abs()], 2)
        o1, _ = gru(feat)
        mask = torch.sigmoid(lin2(torch.relu(lin1(torch.cat([o1, mm], 2)))))
        gain = ((mask * mm) @ erb.T).transpose(1, 2)
        est = torch.cat([gain * sm[:, :257], gain * sm[:, 257:]], 1)
        return ref_istft(est).squeeze(1) + 1e-9

    o_ours, o_ref = net(x, y), ref_stage2(x, y)
    res["stage2_inference_ms"] = {"ours": timeit(lambda: net(x, y), 5), "torch_cudnn_restatement": timeit(lambda: ref_stage2(x, y), 5)}
    # the three launches of the inference separately (same buffers, in_norm shifts precomputed)
    import ctypes as C

    from acoustic_echo_cancellation_b200 import _lib
    lib = _lib.load()
    T = A.num_frames(L)
    sm_, sr_ = A.batch_shift(x), A.batch_shift(y)
    feat = A.stage2_features(x, y, erb, shifts=(sm_, sr_))
    est = torch.empty((B, T, 32), device="cuda")
    outb = torch.empty((B, A.out_samples(L)), device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    res["stage2_parts_ms"] = {
        "batch_shift_x2": timeit(lambda: (A.batch_shift(x), A.batch_shift(y))),
        "features": timeit(lambda: A.stage2_features(x, y, erb, shifts=(sm_, sr_))),
        "mask_gru_linears": timeit(lambda: lib.aec_stage2_mask(feat.data_ptr(), C.byref(net._cw), est.data_ptr(), B, T, 32, st)),
        "synth": timeit(lambda: lib.aec_stage2_synth_dev(x.data_ptr(), est.data_ptr(), net.erb.data_ptr(), outb.data_ptr(), B, L, L,
                                                         outb.shape[1], 512, 32, sm_.data_ptr(), st)),
    }
    n_cmp = min(o_ours.shape[1], o_ref.shape[1])
    res["max_abs_diff"]["stage2"] = float((o_ours[:, :n_cmp] - o_ref[:, :n_cmp]).abs().max())
    res["stage2_out_scale"] = float(o_ref.abs().max())
    print(json.dumps(res, indent=1))
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open("gpurun_out/ops_bench.json", "w"), indent=1)


if __name__ == "__main__":
    main()
