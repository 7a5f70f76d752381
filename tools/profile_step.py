"""Smallest program that launches the hot kernel at a bench workload (for ncu captures).
    python tools/profile_step.py [--config 2|3] [--launches 4] [--variant V] [--batch B]"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import acoustic_echo_cancellation_b200 as A  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--config", type=int, default=2)
ap.add_argument("--launches", type=int, default=4)
ap.add_argument("--variant", type=int, default=0)
ap.add_argument("--batch", type=int, default=0)
ap.add_argument("--algo", type=int, default=-1)
ap.add_argument("--partitions", type=int, default=0)
a = ap.parse_args()
B, P, algo = (1024, 4, 0) if a.config == 2 else (4096, 16, 1)
if a.batch:
    B = a.batch
if a.algo >= 0:
    algo = a.algo
if a.partitions:
    P = a.partitions
L = 160000
g = torch.Generator(device="cuda").manual_seed(1)
far = 0.1 * torch.randn(B, L, device="cuda", generator=g)
mic = 0.5 * torch.roll(far, 37, dims=1) + 0.001 * torch.randn(B, L, device="cuda", generator=g)
out = torch.empty_like(far)
cfg = A.Stage1Config(partitions=P, algo=algo, erle_skip_hops=125, variant=a.variant)
for _ in range(a.launches):
    A.stage1_aec(far, mic, cfg, out=out, return_erle=True)
torch.cuda.synchronize()
print("ok", float(out.abs().mean()))
