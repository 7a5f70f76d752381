"""Smallest program that launches the fused stage-1 + feature kernel and the Stage-2 kernels (for ncu captures).
    python tools/profile_fused.py [--batch 1024]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import acoustic_echo_cancellation_b200 as A  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=1024)
a = ap.parse_args()
B, L = a.batch, 160000
g = torch.Generator(device="cuda").manual_seed(1)
far = 0.1 * torch.randn(B, L, device="cuda", generator=g)
mic = 0.5 * torch.roll(far, 37, dims=1) + 0.001 * torch.randn(B, L, device="cuda", generator=g)
erb = torch.from_numpy(A.erb_filterbank()).float().cuda()
cfg = A.Stage1Config(partitions=4, erle_skip_hops=125)
for _ in range(4):
    err, feat = A.stage1_aec_features(far, mic, erb, cfg)
torch.manual_seed(0)
gru = torch.nn.GRU(64, 32, batch_first=True).cuda().eval()
lin1, lin2 = torch.nn.Linear(64, 32).cuda().eval(), torch.nn.Linear(32, 32).cuda().eval()
sd = {"gru1.weight_ih_l0": gru.weight_ih_l0, "gru1.weight_hh_l0": gru.weight_hh_l0, "gru1.bias_ih_l0": gru.bias_ih_l0,
      "gru1.bias_hh_l0": gru.bias_hh_l0, "linear1.weight": lin1.weight, "linear1.bias": lin1.bias,
      "linear2.weight": lin2.weight, "linear2.bias": lin2.bias}
net = A.LittleNetInference({k: v.detach() for k, v in sd.items()}, erb)
for _ in range(3):
    out = net(err[:128], far[:128])
torch.cuda.synchronize()
print("ok", float(feat.abs().mean()), float(out.abs().mean()))
