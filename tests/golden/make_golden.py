"""Generate golden vectors by IMPORTING THE REFERENCE (run in the build container only).

    python tests/golden/make_golden.py            # writes tests/golden/reference_ops.npz
    python tests/golden/make_golden.py stage2     # writes tests/golden/reference_stage2.npz (Little_net, seeded weights)
    python tests/golden/make_golden.py full       # writes tests/golden/reference_full_size.npz (10 s summaries)

The reference (SZU-Speech/Acoustic-Echo-Cancellation, mounted read-only at
/root/reference) has no tests and no fixtures, so the golden vectors are outputs
of its own operators on seeded inputs:

  ConvSTFT / ConviSTFT          Stage2_lhm/scripts/network/attention_ccrn.py:28-101
  EquivalentRectangularBandwidth Stage2_lhm/scripts/network/ERB.py:10-71
  feature front end              Stage2_lhm/scripts/network/ERB.py:254-290 (restated line by
                                 line with the reference's own modules: the forward() of
                                 Little_net cannot be cut before the GRU)
  countFrames                    Stage2_lhm/scripts/utils/tools.py:30-32

/root/reference does not exist on the GPU box; the tests read only the .npz.
The reference contains no FDAF, so there is no golden vector for the recurrence
(parity unpinned -- see DESIGN.md).
"""
import os
import sys

import numpy as np
import torch

REF = "/root/reference/Stage2_lhm/scripts"
sys.path.insert(0, REF)

from network.attention_ccrn import ConvSTFT, ConviSTFT  # noqa: E402
from network.ERB import EquivalentRectangularBandwidth  # noqa: E402
from utils.tools import countFrames  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from fullsize import full_size_inputs, full_size_positions  # noqa: E402


def main():
    torch.set_num_threads(1)
    torch.manual_seed(0)
    rng = np.random.default_rng(20261018)
    out = {}

    # --- STFT / iSTFT at the live configuration (16 kHz, 512/256, hann) -------------
    stft = ConvSTFT(512, 256, 512, "hann", "complex", fix=True)
    istft = ConviSTFT(512, 256, 512, "hann", "complex", fix=True)
    lengths = [4096, 4095, 4097, 2560, 300]          # ragged: H | L, H !| L, tiny
    for i, n in enumerate(lengths):
        x = (0.3 * rng.standard_normal((2, n))).astype(np.float32)
        with torch.no_grad():
            s = stft(torch.from_numpy(x))
            y = istft(s)
        out[f"x512_{i}"] = x
        out[f"stft512_{i}"] = s.numpy()
        out[f"istft512_{i}"] = y.numpy()
    # iSTFT of a spectrum that is NOT the STFT of a signal (non-zero imag at DC/Nyquist)
    s = (0.5 * rng.standard_normal((2, 514, 9))).astype(np.float32)
    with torch.no_grad():
        out["spec_free"] = s
        out["istft_free"] = istft(torch.from_numpy(s)).numpy()

    # --- the 48 kHz configuration (frame 1024, hop 512) -------------------------------
    stft1k = ConvSTFT(1024, 512, 1024, "hann", "complex", fix=True)
    istft1k = ConviSTFT(1024, 512, 1024, "hann", "complex", fix=True)
    x = (0.3 * rng.standard_normal((2, 6144))).astype(np.float32)
    with torch.no_grad():
        s = stft1k(torch.from_numpy(x))
        y = istft1k(s)
    out["x1024"] = x
    out["stft1024"] = s.numpy()
    out["istft1024"] = y.numpy()

    # --- frame counts -------------------------------------------------------------------
    ls = np.array([159999, 160000, 160001, 480000, 4096, 300, 256, 255, 1], dtype=np.int64)
    frames = []
    for n in ls:
        with torch.no_grad():
            frames.append(stft(torch.zeros(1, int(n))).shape[-1])
    out["frame_count_L"] = ls
    out["frame_count_T"] = np.array(frames, dtype=np.int64)
    out["countFrames_ref"] = np.array([countFrames(int(n), 512, 256) for n in ls], dtype=np.int64)

    # --- ERB filterbank and the Stage-2 feature front end -------------------------------
    erb = EquivalentRectangularBandwidth(257, 16000, 32, 0, 8000).filters
    out["erb"] = erb                                          # float64 [257, 32]
    mic = (0.2 * rng.standard_normal((3, 8192)) + 0.01).astype(np.float32)
    ref = (0.2 * rng.standard_normal((3, 8192)) - 0.02).astype(np.float32)
    with torch.no_grad():
        erb_t = torch.from_numpy(erb).float()                 # train1.py:148 casts to float32
        m = torch.from_numpy(mic)
        r = torch.from_numpy(ref)
        m = m - (torch.mean(m) / torch.std(m))                # ERB.py:254
        r = r - (torch.mean(r) / torch.std(r))                # ERB.py:255
        ms, rs = stft(m), stft(r)                             # ERB.py:263-264
        k = 257
        mmag = torch.sqrt(ms[:, :k] ** 2 + ms[:, k:] ** 2 + 1e-9).transpose(1, 2)   # ERB.py:277
        rmag = torch.sqrt(rs[:, :k] ** 2 + rs[:, k:] ** 2 + 1e-9).transpose(1, 2)   # ERB.py:278
        merb, rerb = mmag @ erb_t, rmag @ erb_t               # ERB.py:282-283
        feat = torch.cat([merb, torch.abs(merb - rerb)], dim=2)  # ERB.py:287-290
    out["feat_mic"] = mic
    out["feat_ref"] = ref
    out["feat"] = feat.numpy()

    path = os.path.join(HERE, "reference_ops.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;", len(out), "arrays")


def stage2_golden():
    """Little_net inference (Stage2_lhm/scripts/network/ERB.py:203-334) with seeded random weights:
    the reference module itself is run on CPU; weights, inputs and `out_wav` go to
    tests/golden/reference_stage2.npz."""
    from network.ERB import Little_net

    torch.manual_seed(1234)
    conf = {"win_size": 512, "hop_size": 256}
    net = Little_net(conf, 32).eval()
    with torch.no_grad():
        for p_ in net.parameters():            # lively, well-conditioned random weights
            p_.copy_(0.3 * torch.randn_like(p_))
    rng = np.random.default_rng(777)
    mic = (0.2 * rng.standard_normal((3, 6144)) + 0.01).astype(np.float32)
    ref = (0.2 * rng.standard_normal((3, 6144)) - 0.02).astype(np.float32)
    near = (0.1 * rng.standard_normal((3, 6144))).astype(np.float32)
    erb = EquivalentRectangularBandwidth(257, 16000, 32, 0, 8000).filters
    with torch.no_grad():
        out_wav, loss = net(torch.from_numpy(mic), torch.from_numpy(ref), torch.from_numpy(near),
                            torch.from_numpy(erb).float())
    out = {"mic": mic, "ref": ref, "erb": erb.astype(np.float32), "out_wav": out_wav.numpy()}
    for k, v in net.state_dict().items():
        if k.startswith(("gru1.", "linear1.", "linear2.")):
            out["w_" + k.replace(".", "_")] = v.numpy()
    path = os.path.join(HERE, "reference_stage2.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;", sorted(out))


def full_size_golden():
    """BASELINE.json's utterance length (10 s, 16 kHz) through the reference's own operators and its live model:
    the outputs are too large to commit, so tests/golden/reference_full_size.npz keeps 256 single values at fixed
    positions plus per-channel / per-frame sums of each output (float64) -- enough to catch a wrong frame count,
    layout, window, normalisation or band anywhere in the 626 frames."""
    from network.ERB import Little_net

    torch.set_num_threads(4)
    mic, ref = full_size_inputs()
    stft = ConvSTFT(512, 256, 512, "hann", "complex", fix=True)
    istft = ConviSTFT(512, 256, 512, "hann", "complex", fix=True)
    erb = EquivalentRectangularBandwidth(257, 16000, 32, 0, 8000).filters
    out = {}
    with torch.no_grad():
        s = stft(torch.from_numpy(mic))                                   # [2, 514, 626]
        y = istft(s)                                                      # [2, 1, 160000]
        s_np, y_np = s.numpy(), y.numpy()
        out["stft_shape"], out["istft_shape"] = np.array(s_np.shape), np.array(y_np.shape)
        out["stft_vals"] = s_np.ravel()[full_size_positions(s_np.shape)]
        out["stft_sum_t"] = s_np.astype(np.float64).sum(axis=2)           # [2, 514]
        out["stft_pow_c"] = (s_np.astype(np.float64) ** 2).sum(axis=1)    # [2, 626]
        out["istft_vals"] = y_np.ravel()[full_size_positions(y_np.shape)]
        out["istft_pow"] = (y_np.astype(np.float64) ** 2).sum(axis=(1, 2))
        # feature front end (ERB.py:254-290), restated with the reference's modules as in main()
        m_t, r_t = torch.from_numpy(mic), torch.from_numpy(ref)
        m_t = m_t - m_t.mean() / m_t.std()
        r_t = r_t - r_t.mean() / r_t.std()
        feats = []
        for v in (m_t, r_t):
            sp = stft(v)
            mag = torch.sqrt(sp[:, :257] ** 2 + sp[:, 257:] ** 2 + 1e-9).transpose(1, 2)
            feats.append(mag @ torch.from_numpy(erb).float())
        f_np = torch.cat([feats[0], (feats[0] - feats[1]).abs()], 2).numpy()          # [2, 626, 64]
        out["feat_shape"] = np.array(f_np.shape)
        out["feat_vals"] = f_np.ravel()[full_size_positions(f_np.shape)]
        out["feat_sum_t"] = f_np.astype(np.float64).sum(axis=1)                         # [2, 64]
        # the live model with the seeded weights of reference_stage2.npz
        g = np.load(os.path.join(HERE, "reference_stage2.npz"))
        net = Little_net({"win_size": 512, "hop_size": 256}, 32).eval()
        sd = net.state_dict()
        for k in list(sd):
            key = "w_" + k.replace(".", "_")
            if key in g.files:
                sd[k] = torch.from_numpy(g[key])
        net.load_state_dict(sd)
        near = torch.zeros_like(m_t)
        o, _ = net(torch.from_numpy(mic), torch.from_numpy(ref), near, torch.from_numpy(erb).float())
        o_np = o.numpy()
        out["net_shape"] = np.array(o_np.shape)
        out["net_vals"] = o_np.ravel()[full_size_positions(o_np.shape)]
        out["net_pow"] = (o_np.astype(np.float64) ** 2).sum(axis=-1).reshape(-1)
    path = os.path.join(HERE, "reference_full_size.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "stage2":
        stage2_golden()
    elif len(sys.argv) > 1 and sys.argv[1] == "full":
        full_size_golden()
    else:
        main()
