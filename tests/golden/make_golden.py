"""Generate golden vectors by IMPORTING THE REFERENCE (run in the build container only).

    python tests/golden/make_golden.py            # writes tests/golden/reference_ops.npz

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


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "stage2":
        stage2_golden()
    else:
        main()
