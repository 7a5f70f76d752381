"""Developer experiment: why does the PCM16 host pipeline sit at 86 % of the raw PCIe floor on one GPU?
Emulates the slice pipeline with torch streams in three shapes and times one config-2 step (1024 x 10 s):
  slots   : one stream per slice slot carrying H2D -> convert -> kernel -> D2H (the library's shape)
  lanes   : one upload stream, one compute stream, one download stream, chained by events
  copies  : `lanes` without any kernel (sliced copies only)"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import acoustic_echo_cancellation_b200 as A  # noqa: E402

B, L = 1024, 160000
cfg = A.Stage1Config()
hf = torch.from_numpy(A.pinned_empty((B, L), np.int16)); hm = torch.from_numpy(A.pinned_empty((B, L), np.int16))
hf.random_(-3000, 3000); hm.random_(-3000, 3000)
he = torch.from_numpy(A.pinned_empty((B, L), np.float32))


def run(shape, sl, nslots, kernels=True):
    d16f = [torch.empty(sl, L, dtype=torch.int16, device="cuda") for _ in range(nslots)]
    d16m = [torch.empty(sl, L, dtype=torch.int16, device="cuda") for _ in range(nslots)]
    df = [torch.empty(sl, L, device="cuda") for _ in range(nslots)]
    dm = [torch.empty(sl, L, device="cuda") for _ in range(nslots)]
    de = [torch.empty(sl, L, device="cuda") for _ in range(nslots)]
    streams = [torch.cuda.Stream() for _ in range(nslots)]
    up, comp, down = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()
    free_ev = [None] * nslots

    def step():
        for it, off in enumerate(range(0, B, sl)):
            k = it % nslots
            nb = min(sl, B - off)
            if shape == "slots":
                s = streams[k]
                s.synchronize()
                with torch.cuda.stream(s):
                    d16f[k][:nb].copy_(hf[off:off + nb], non_blocking=True)
                    d16m[k][:nb].copy_(hm[off:off + nb], non_blocking=True)
                    if kernels:
                        torch.mul(d16f[k][:nb], 1.0 / 32768.0, out=df[k][:nb])
                        torch.mul(d16m[k][:nb], 1.0 / 32768.0, out=dm[k][:nb])
                        A.stage1_aec(df[k][:nb], dm[k][:nb], cfg, out=de[k][:nb])
                    he[off:off + nb].copy_(de[k][:nb], non_blocking=True)
            else:
                if free_ev[k] is not None:
                    up.wait_event(free_ev[k])
                with torch.cuda.stream(up):
                    d16f[k][:nb].copy_(hf[off:off + nb], non_blocking=True)
                    d16m[k][:nb].copy_(hm[off:off + nb], non_blocking=True)
                    e_up = torch.cuda.Event(); e_up.record()
                comp.wait_event(e_up)
                with torch.cuda.stream(comp):
                    if kernels:
                        torch.mul(d16f[k][:nb], 1.0 / 32768.0, out=df[k][:nb])
                        torch.mul(d16m[k][:nb], 1.0 / 32768.0, out=dm[k][:nb])
                        A.stage1_aec(df[k][:nb], dm[k][:nb], cfg, out=de[k][:nb])
                    e_c = torch.cuda.Event(); e_c.record()
                down.wait_event(e_c)
                with torch.cuda.stream(down):
                    he[off:off + nb].copy_(de[k][:nb], non_blocking=True)
                    free_ev[k] = torch.cuda.Event(); free_ev[k].record()
        torch.cuda.synchronize()

    step()
    t0 = time.perf_counter()
    for _ in range(4):
        step()
    return (time.perf_counter() - t0) / 4 * 1e3


for shape, sl, ns, kern in [("slots", 128, 4, True), ("lanes", 128, 4, True), ("lanes", 64, 8, True), ("lanes", 128, 4, False),
                            ("slots", 128, 4, False), ("lanes", 256, 3, True), ("lanes", 32, 8, True)]:
    print(f"{shape:6s} slice {sl:4d} slots {ns} kernels {kern}: {run(shape, sl, ns, kern):.2f} ms")
