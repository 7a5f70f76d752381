"""Device-timed throughput of the overlap-save kernels (algo 2 / 3) next to the STFT-domain NLMS kernel.
    python tools/ols_time.py [B ...]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import acoustic_echo_cancellation_b200 as A  # noqa: E402

FRAME = int(os.environ.get("OLS_FRAME", "512"))
for B in [int(a) for a in sys.argv[1:]] or [1024, 4144]:
    L = 160000 if FRAME == 512 else 480000          # 10 s at 16 / 48 kHz
    g = torch.Generator(device="cuda").manual_seed(1)
    far = 0.1 * torch.randn(B, L, device="cuda", generator=g)
    mic = 0.5 * torch.roll(far, 37, dims=1) + 0.001 * torch.randn(B, L, device="cuda", generator=g)
    out = torch.empty_like(far)
    for algo, P, var in (((0, 4, 0), (2, 1, 0), (2, 2, 0), (2, 4, 0), (3, 4, 0), (0, 8, 0), (2, 8, 0), (3, 8, 0), (1, 16, 0), (2, 16, 0), (3, 16, 0)) if FRAME == 512 else
                         ((0, 8, 0), (1, 8, 0), (2, 8, 0), (3, 8, 0), (0, 4, 0), (2, 4, 0), (3, 4, 0))):
        cfg = A.Stage1Config(frame=FRAME, partitions=P, algo=algo, erle_skip_hops=125, variant=var)
        for _ in range(3):
            A.stage1_aec(far, mic, cfg, out=out, return_erle=True)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5):
            e, erle = A.stage1_aec(far, mic, cfg, out=out, return_erle=True)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 5
        print(f"frame={FRAME} B={B} algo={algo} P={P} variant={var}: {ms:.3f} ms  {B * 10 / ms / 1e3:.2f} M audio-s/s  erle {float(erle.mean()):.1f} dB", flush=True)
    del far, mic, out
