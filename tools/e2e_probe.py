"""PCIe / host-pipeline probe: raw pinned H2D, D2H and concurrent rates, then HostPipeline.run at
several slice sizes (config-2 workload)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import acoustic_echo_cancellation_b200 as A
B, L = 1024, 160000
h = torch.empty(B, L, pin_memory=True); d = torch.empty(B, L, device="cuda")
h2 = torch.empty(B, L, pin_memory=True); d2 = torch.empty(B, L, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(fn, n=5):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n
gb = B * L * 4 / 1e9
print("H2D GB/s", gb / t(lambda: d.copy_(h, non_blocking=True)))
print("D2H GB/s", gb / t(lambda: h.copy_(d, non_blocking=True)))
def both():
    with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
print("concurrent H2D+D2H GB/s each", gb / t(both))
hf, hm, he = A.pinned_empty((B, L)), A.pinned_empty((B, L)), A.pinned_empty((B, L))
hf[:] = np.random.default_rng(0).standard_normal((B, L)).astype(np.float32) * 0.1; hm[:] = hf * 0.5
cfg = A.Stage1Config()
for sl in (32, 64, 128, 256, 512):
    pipe = A.HostPipeline(sl, L)
    dt = t(lambda: pipe.run(hf, hm, cfg, err=he), 4)
    print(f"slice {sl}: {dt*1e3:.2f} ms/step  {B*10/dt/1e3:.1f}k audio-s/s  ({3*gb/dt:.1f} GB/s total)")
    pipe.close()
h16f, h16m = A.pinned_empty((B, L), dtype=np.int16), A.pinned_empty((B, L), dtype=np.int16)
h16f[:] = np.clip(np.rint(hf * 32768.0), -32768, 32767).astype(np.int16); h16m[:] = np.clip(np.rint(hm * 32768.0), -32768, 32767).astype(np.int16)
for sl in (16, 32, 64, 128):
    pipe = A.HostPipeline(sl, L)
    dt = t(lambda: pipe.run(h16f, h16m, cfg, err=he), 4)
    print(f"pcm16 slice {sl}: {dt*1e3:.2f} ms/step  {B*10/dt/1e3:.1f}k audio-s/s")
    pipe.close()
