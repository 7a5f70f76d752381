import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import acoustic_echo_cancellation_b200 as A
B, L = 128, 160000
x = 0.1 * torch.randn(B, L, device="cuda"); y = 0.1 * torch.randn(B, L, device="cuda")
erb = torch.from_numpy(A.erb_filterbank()).float().cuda()
for _ in range(3):
    f = A.stage2_features(x, y, erb, in_norm=False)
torch.cuda.synchronize(); print("ok", float(f.mean()))
