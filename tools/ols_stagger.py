"""Overlap-save kernels: does a start-up skew between co-resident utterances (aec_cfg.stagger_ns) help?
    python tools/ols_stagger.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import acoustic_echo_cancellation_b200 as A  # noqa: E402

L = 160000
for B in (1024, 4144):
    g = torch.Generator(device="cuda").manual_seed(1)
    far = 0.1 * torch.randn(B, L, device="cuda", generator=g)
    mic = 0.5 * torch.roll(far, 37, dims=1) + 0.001 * torch.randn(B, L, device="cuda", generator=g)
    out = torch.empty_like(far)
    ref = None
    for algo, P in ((2, 4), (3, 4), (3, 16)):
        if P == 16 and B > 2048:
            continue
        for st in (0, 150, 300, 450, 700, 1000, 2000):
            cfg = A.Stage1Config(partitions=P, algo=algo, erle_skip_hops=125, stagger_ns=st)
            for _ in range(2):
                A.stage1_aec(far, mic, cfg, out=out, return_erle=True)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(5):
                e, erle = A.stage1_aec(far, mic, cfg, out=out, return_erle=True)
            b.record()
            torch.cuda.synchronize()
            chk = float(out.double().abs().sum())
            if st == 0:
                ref = chk
            print(f"B={B} algo={algo} P={P} stagger_ns={st}: {a.elapsed_time(b) / 5:.3f} ms  same_output={chk == ref}", flush=True)
