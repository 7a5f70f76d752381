"""Small runs of every kernel family for compute-sanitizer (memcheck / racecheck / synccheck).
    compute-sanitizer --tool racecheck python tools/sanitize_step.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import acoustic_echo_cancellation_b200 as A  # noqa: E402

torch.manual_seed(0)
for frame, P, algo, echo, L in [(512, 4, 0, False, 6000), (512, 4, 0, True, 6000), (512, 4, 1, False, 5000), (512, 8, 0, False, 9000),
                                (512, 16, 1, False, 9000), (512, 16, 1, True, 5000), (512, 16, 0, False, 7000),
                                (1024, 8, 0, False, 12000)]:
    B = 3
    far = 0.1 * torch.randn(B, L, device="cuda")
    mic = 0.5 * torch.roll(far, 37, dims=1)
    ns = torch.tensor([L, L - 301, 700], device="cuda", dtype=torch.int64)
    cfg = A.Stage1Config(frame=frame, partitions=P, algo=algo)
    res = A.stage1_aec(far, mic, cfg, n_samples=ns, return_echo=echo, return_erle=True)
    torch.cuda.synchronize()
    print(frame, P, algo, echo, float(res[0].abs().mean()))
x = torch.randn(2, 8192, device="cuda")
s = A.ConvSTFT(512, 256, 512, "hann", "complex")(x)
y = A.ConviSTFT(512, 256, 512, "hann", "complex")(s)
torch.cuda.synchronize()
print("ok")
