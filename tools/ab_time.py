"""Developer tool: time alternative builds of the library (ab/lib_*.so) on the same inputs.

    python tools/ab_time.py v1 v2 --batches 888,1024,4144 [--partitions 4 --algo 0]
Each build runs in its own process (AEC_B200_LIB selects the shared object)."""
import argparse
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r"""
import json, sys, torch
sys.path.insert(0, %r)
import acoustic_echo_cancellation_b200 as A
P, algo, frame, variant = %d, %d, %d, %d
for B in %r:
    L = 160000 * frame // 512
    g = torch.Generator(device="cuda").manual_seed(1)
    far = 0.1 * torch.randn(B, L, device="cuda", generator=g)
    mic = 0.5 * torch.roll(far, 37, dims=1) + 0.001 * torch.randn(B, L, device="cuda", generator=g)
    out = torch.empty_like(far)
    cfg = A.Stage1Config(frame=frame, partitions=P, algo=algo, variant=variant)
    for _ in range(3):
        A.stage1_aec(far, mic, cfg, out=out)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(8)]
    ev[0].record()
    for i in range(7):
        A.stage1_aec(far, mic, cfg, out=out)
        ev[i + 1].record()
    torch.cuda.synchronize()
    ms = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(7))
    print(json.dumps({"B": B, "ms_best": round(ms[0], 4), "ms_med": round(ms[3], 4), "sum": float(out.double().sum())}), flush=True)
"""


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("names", nargs="+")
    ap.add_argument("--batches", default="888,1024,4144")
    ap.add_argument("--partitions", type=int, default=4)
    ap.add_argument("--algo", type=int, default=0)
    ap.add_argument("--frame", type=int, default=512)
    ap.add_argument("--variant", type=int, default=0)
    args = ap.parse_args()
    batches = [int(x) for x in args.batches.split(",")]
    for name in args.names:
        lib = os.path.join(ROOT, "acoustic_echo_cancellation_b200", "libaec_b200.so") if name == "default" \
            else os.path.join(ROOT, "ab", f"lib_{name}.so")
        env = dict(os.environ, AEC_B200_LIB=lib)
        code = CHILD % (ROOT, args.partitions, args.algo, args.frame, args.variant, batches)
        r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
        for line in r.stdout.splitlines():
            print(name, line, flush=True)
        if r.returncode:
            print(name, "FAILED", r.stderr[-500:], flush=True)


if __name__ == "__main__":
    main()
