import os, sys, time
import numpy as np, torch
sys.path.insert(0, "/root/repo")
import acoustic_echo_cancellation_b200 as A
B, L = 1024, 160000
hf, hm, he = A.pinned_empty((B, L)), A.pinned_empty((B, L)), A.pinned_empty((B, L))
hf[:] = np.random.default_rng(0).standard_normal((B, L)).astype(np.float32) * 0.1; hm[:] = hf * 0.5
h16f, h16m = A.pinned_empty((B, L), dtype=np.int16), A.pinned_empty((B, L), dtype=np.int16)
h16f[:] = np.clip(np.rint(hf * 32768.0), -32768, 32767).astype(np.int16); h16m[:] = np.clip(np.rint(hm * 32768.0), -32768, 32767).astype(np.int16)
cfg = A.Stage1Config()
def t(fn, n=6):
    fn(); fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n
for sl in (64, 128):
    pipe = A.HostPipeline(sl, L)
    a = t(lambda: pipe.run(hf, hm, cfg, err=he)); ref = he.copy()
    b = t(lambda: pipe.run(h16f, h16m, cfg, err=he))
    print(os.environ.get("AEC_B200_LIB", "default").split("/")[-1], "slice", sl, "f32 %.2f ms  pcm16 %.2f ms" % (a * 1e3, b * 1e3), float(np.abs(ref).sum()))
    pipe.close()
