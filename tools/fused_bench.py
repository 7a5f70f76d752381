"""Stage 1 -> Stage-2 feature front end: the fused kernel (aec_stage1_run_features) against the two-launch pipeline
(aec_stage1_run, then aec_features on the stored error signal and the far end), device-timed on the config-2 batch.

    python tools/fused_bench.py [--batch 1024]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import acoustic_echo_cancellation_b200 as A  # noqa: E402


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=1024)
    ap.add_argument("--algo", type=int, default=0)
    a = ap.parse_args()
    B, L = a.batch, 160000
    g = torch.Generator(device="cuda").manual_seed(1)
    far = 0.1 * torch.randn(B, L, device="cuda", generator=g)
    mic = 0.5 * torch.roll(far, 37, dims=1) + 0.001 * torch.randn(B, L, device="cuda", generator=g)
    erb = torch.from_numpy(A.erb_filterbank()).float().cuda()
    cfg = A.Stage1Config(partitions=4, algo=a.algo, erle_skip_hops=125)
    err = torch.empty_like(far)
    res = {"batch": B, "samples": L, "algo": "nlms" if a.algo == 0 else "kalman"}
    res["stage1_ms"] = timeit(lambda: A.stage1_aec(far, mic, cfg, out=err))
    res["features_standalone_ms"] = timeit(lambda: A.stage2_features(err, far, erb, in_norm=False))
    res["two_launch_pipeline_ms"] = timeit(lambda: (A.stage1_aec(far, mic, cfg, out=err),
                                                     A.stage2_features(err, far, erb, in_norm=False)))
    res["fused_ms"] = timeit(lambda: A.stage1_aec_features(far, mic, erb, cfg, out=err))
    e2, f2 = A.stage1_aec_features(far, mic, erb, cfg)
    want = A.stage2_features(A.stage1_aec(far, mic, cfg), far, erb, in_norm=False)
    res["max_abs_diff_features"] = float((f2 - want).abs().max())
    res["features_scale"] = float(want.abs().max())
    res["speedup_vs_two_launches"] = res["two_launch_pipeline_ms"] / res["fused_ms"]
    res["hbm_bytes_two_launches"] = 3 * 4 * B * L + 2 * 4 * B * L + 4 * B * A.num_frames(L) * 64
    res["hbm_bytes_fused"] = 3 * 4 * B * L + 4 * B * A.num_frames(L) * 64
    print(json.dumps(res, indent=1))
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open(f"gpurun_out/fused_bench_{B}_{a.algo}.json", "w"), indent=1)


if __name__ == "__main__":
    main()
