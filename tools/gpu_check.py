"""Developer check run on the GPU box: parity of every kernel family against the numpy oracle on
small cases, then a timing sweep over tuning variants.  Writes gpurun_out/gpu_check.json.

    python tools/gpu_check.py [--no-sweep] [--variants 207,206,...]
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import acoustic_echo_cancellation_b200 as A  # noqa: E402
from acoustic_echo_cancellation_b200 import synth  # noqa: E402
from oracle import aec_oracle as O  # noqa: E402


def parity_case(name, P, algo, L, B=3, ragged=False, echo=True, variant=0, unaligned=False, double_talk=False, frame=512):
    hop = frame // 2
    d = synth.make_batch(0, B, L, sample_rate=16000 * frame // 512, rir_len=min(P * hop, 4096), double_talk=double_talk)
    far, mic = d["far"], d["mic"]
    ns = None
    if ragged:
        ns = np.array([L, L - 1, max(L - 777, 1)][:B] + [L // 2] * max(B - 3, 0), dtype=np.int64)
    cfg_o = O.AecConfig(frame=frame, partitions=P, algo=algo, delta=1e-6 * frame)
    skip_hops = 8
    ref = O.stage1(far, mic, cfg_o, n_samples=ns, erle_skip=skip_hops * hop)
    cfg = A.Stage1Config(frame=frame, partitions=P, algo=algo, erle_skip_hops=skip_hops, variant=variant)
    if unaligned:
        buf_f = torch.zeros(B, L + 3, device="cuda")
        buf_m = torch.zeros(B, L + 3, device="cuda")
        buf_f[:, 1:L + 1] = torch.from_numpy(far).cuda()
        buf_m[:, 1:L + 1] = torch.from_numpy(mic).cuda()
        tf, tm = buf_f[:, 1:L + 1], buf_m[:, 1:L + 1]
    else:
        tf, tm = torch.from_numpy(far).cuda(), torch.from_numpy(mic).cuda()
    tn = torch.from_numpy(ns).cuda() if ns is not None else None
    res = A.stage1_aec(tf, tm, cfg, n_samples=tn, return_echo=echo, return_erle=True)
    torch.cuda.synchronize()
    if echo:
        err, ec, erle = res
    else:
        err, erle = res
        ec = None
    err = err.cpu().numpy()
    lo = ref["err"].shape[1]
    out = {"case": name, "frame": frame, "P": P, "algo": algo, "L": L, "B": B, "variant": variant,
           "max_abs_err": float(np.abs(err[:, :lo] - ref["err"]).max()) if lo else 0.0,
           "tail_zero": bool((err[:, lo:] == 0).all()),
           "erle_diff_db": float(np.abs(erle.cpu().numpy() - ref["erle_db"]).max()),
           "erle_db": [float(x) for x in ref["erle_db"]]}
    if ns is not None:
        for b in range(B):
            m = O.n_frames(int(ns[b]), frame, hop) - 1
            m = max(m, 0) * hop
            out["tail_zero"] = out["tail_zero"] and bool((err[b, m:] == 0).all())
    if ec is not None:
        out["max_abs_echo"] = float(np.abs(ec.cpu().numpy()[:, :lo] - ref["echo"]).max()) if lo else 0.0
    print(json.dumps(out), flush=True)
    return out


def spectral_cases():
    outs = []
    rng = np.random.default_rng(5)
    for L in (4096, 4095, 4097, 300, 16000):
        x = (0.3 * rng.standard_normal((3, L))).astype(np.float32)
        s_ref = O.stft(x)
        tx = torch.from_numpy(x).cuda()
        s = A.ConvSTFT(512, 256, 512, "hann", "complex")(tx)
        y = A.ConviSTFT(512, 256, 512, "hann", "complex")(s)
        y_ref = O.istft(s_ref)
        torch.cuda.synchronize()
        o = {"case": f"stft/istft L={L}", "stft_max": float(np.abs(s.cpu().numpy() - s_ref).max()),
             "istft_max": float(np.abs(y.cpu().numpy() - y_ref).max()) if y_ref.size else 0.0,
             "shape_ok": list(s.shape) == list(s_ref.shape) and list(y.shape) == list(y_ref.shape)}
        print(json.dumps(o), flush=True)
        outs.append(o)
    # free spectrum (non-zero imag at DC/Nyquist)
    sp = (0.5 * rng.standard_normal((2, 514, 37))).astype(np.float32)
    y = A.ConviSTFT(512, 256, 512, "hann", "complex")(torch.from_numpy(sp).cuda()).cpu().numpy()
    o = {"case": "istft free spectrum", "istft_max": float(np.abs(y - O.istft(sp.astype(np.float64))).max())}
    print(json.dumps(o), flush=True)
    outs.append(o)
    erb = O.erb_filterbank()
    mic = (0.2 * rng.standard_normal((3, 8192)) + 0.01).astype(np.float32)
    ref = (0.2 * rng.standard_normal((3, 8192)) - 0.02).astype(np.float32)
    f_ref = O.stage2_features(mic, ref, erb)
    f = A.stage2_features(torch.from_numpy(mic).cuda(), torch.from_numpy(ref).cuda(),
                          torch.from_numpy(erb).float().cuda()).cpu().numpy()
    o = {"case": "features", "max": float(np.abs(f - f_ref).max()), "scale": float(np.abs(f_ref).max())}
    print(json.dumps(o), flush=True)
    outs.append(o)
    return outs


def time_variant(variant, P=4, algo=0, B=1024, L=160000, iters=5, echo=False, stagger=0, frame=512):
    g = torch.Generator(device="cuda").manual_seed(1)
    far = 0.1 * torch.randn(B, L, device="cuda", generator=g)
    mic = 0.5 * torch.roll(far, 37, dims=1) + 0.001 * torch.randn(B, L, device="cuda", generator=g)
    out = torch.empty_like(far)
    cfg = A.Stage1Config(frame=frame, partitions=P, algo=algo, variant=variant, stagger_ns=stagger)
    try:
        for _ in range(2):
            A.stage1_aec(far, mic, cfg, out=out, return_echo=echo)
        torch.cuda.synchronize()
    except A.AecError as e:
        return {"variant": variant, "error": str(e)}
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(iters + 1)]
    ev[0].record()
    for i in range(iters):
        A.stage1_aec(far, mic, cfg, out=out, return_echo=echo)
        ev[i + 1].record()
    torch.cuda.synchronize()
    ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(iters)]
    best = min(ms)
    audio_s = B * L / (16000.0 * frame / 512)
    o = {"variant": variant, "frame": frame, "stagger": stagger, "P": P, "algo": algo, "B": B, "echo": echo, "ms_best": best, "ms_med": float(np.median(ms)),
         "audio_s_per_s": audio_s / (best * 1e-3), "finite": bool(torch.isfinite(out).all())}
    print(json.dumps(o), flush=True)
    return o


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--no-sweep", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--variants", default="2128,2168,4128,4096,1255,1200")
    ap.add_argument("--stagger", default="0")
    ap.add_argument("--batches", default="1024")
    args = ap.parse_args()
    res = {"parity": [], "spectral": [], "sweep": []}
    print("device", torch.cuda.get_device_name(0), flush=True)
    t0 = time.time()
    if not args.no_parity:
        res["parity"].append(parity_case("nlms P4", 4, 0, 16000))
        res["parity"].append(parity_case("nlms P4 ragged", 4, 0, 16000 + 123, ragged=True))
        res["parity"].append(parity_case("nlms P4 unaligned", 4, 0, 16000 + 1, unaligned=True))
        res["parity"].append(parity_case("nlms P4 noecho", 4, 0, 16000, echo=False))
        res["parity"].append(parity_case("nlms P4 10s", 4, 0, 160000, B=2))
        res["parity"].append(parity_case("kalman P4", 4, 1, 16000))
        res["parity"].append(parity_case("nlms P1", 1, 0, 8000))
        res["parity"].append(parity_case("kalman P2", 2, 1, 8000))
        res["parity"].append(parity_case("nlms P8", 8, 0, 16000))
        res["parity"].append(parity_case("kalman P8", 8, 1, 16000))
        res["parity"].append(parity_case("nlms P16", 16, 0, 16000))
        res["parity"].append(parity_case("kalman P16", 16, 1, 16000))
        res["parity"].append(parity_case("kalman P16 10s", 16, 1, 160000, B=2, echo=False))
        res["parity"].append(parity_case("nlms P4 tiny", 4, 0, 300, echo=False))
        res["parity"].append(parity_case("nlms P4 dt", 4, 0, 32000, double_talk=True))
        for v in (2128, 2168, 4128, 4096, 1255, 1200):
            res["parity"].append(parity_case(f"nlms P4 variant {v}", 4, 0, 16000 + 256, echo=False, variant=v))
        for (pp, aa) in ((16, 1), (16, 0), (8, 1), (8, 0)):
            res["parity"].append(parity_case(f"P{pp} algo{aa} variant 8128", pp, aa, 16000 + 123, echo=True, variant=8128, ragged=True))
        for (pp, aa) in ((8, 0), (8, 1), (4, 0), (4, 1), (2, 1), (1, 0)):
            res["parity"].append(parity_case(f"frame1024 P{pp} algo{aa}", pp, aa, 24000 + 77, echo=True, ragged=True, frame=1024))
        res["parity"].append(parity_case("frame1024 P8 10s dt", 8, 0, 480000, B=2, echo=False, frame=1024, double_talk=True))
        res["parity"].append(parity_case("frame1024 P8 unaligned", 8, 0, 24001, echo=False, frame=1024, unaligned=True))
        res["parity"].append(parity_case("frame1024 P4 tiny", 4, 0, 700, echo=False, frame=1024))
        res["spectral"] = spectral_cases()
    print("parity time", time.time() - t0, flush=True)
    if not args.no_sweep:
        print("fp32 peak TFLOP/s", A.fp32_peak_tflops(), flush=True)
        for v in [int(x) for x in args.variants.split(",") if x]:
            for stg in [int(x) for x in args.stagger.split(",")]:
                for bb in [int(x) for x in args.batches.split(",")]:
                    res["sweep"].append(time_variant(v, stagger=stg, B=bb))
        for v in (4255, 8128):
            res["sweep"].append(time_variant(v, P=16, algo=1, B=2048))
            res["sweep"].append(time_variant(v, P=16, algo=0, B=2048))
        for v in (4168, 4128, 8128):
            res["sweep"].append(time_variant(v, P=8, algo=0, B=1024))
            res["sweep"].append(time_variant(v, P=8, algo=1, B=1024))
        res["sweep"].append(time_variant(0, P=4, algo=1))
        res["sweep"].append(time_variant(0, P=8, algo=0))
        res["sweep"].append(time_variant(0, P=16, algo=1, B=2048))
        res["sweep"].append(time_variant(0, P=4, algo=0, echo=True))
        res["sweep"].append(time_variant(0, P=4, algo=0, B=4096))
        res["sweep"].append(time_variant(0, P=8, algo=0, B=1024, L=480000, frame=1024))
        res["sweep"].append(time_variant(0, P=8, algo=1, B=512, L=480000, frame=1024))
        res["fp32_peak_tflops"] = A.fp32_peak_tflops()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "gpu_check.json"), "w") as f:
        json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
