"""Developer tool: per-phase cycle counts of the fused stage-1 kernel (two-warp frame-512 kernel).

Needs the instrumented build (`make -C acoustic_echo_cancellation_b200/csrc dbg` -> ab/libaec_b200_dbg.so,
compiled with -DAEC_PHASE_TIMING; the product library carries no instrumentation):

    AEC_B200_LIB=ab/libaec_b200_dbg.so python tools/phase_timing.py --batches 148,592,1036
"""
import argparse
import ctypes as C
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("AEC_B200_LIB", os.path.join(ROOT, "ab", "libaec_b200_dbg.so"))

import acoustic_echo_cancellation_b200 as A  # noqa: E402
from acoustic_echo_cancellation_b200 import _lib  # noqa: E402

NAMES = ["loop tail", "wait(bar+mbar)", "phase A", "bar A", "produce", "phase B", "bar B", "phase C", "bar C", "exit"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batches", default="148,592,1036")
    ap.add_argument("--variant", type=int, default=2128)
    ap.add_argument("--partitions", type=int, default=4)
    ap.add_argument("--algo", type=int, default=0)
    ap.add_argument("--samples", type=int, default=160000)
    ap.add_argument("--placement", action="store_true", help="print hardware warp slots of the utterances on one SM")
    args = ap.parse_args()
    lib = _lib.load()
    setter = lib.aec_debug_set_phase_buffer
    setter.argtypes = [C.c_void_p]
    setter.restype = None
    nw = args.variant // 1000 if args.algo < 2 else (2 if args.partitions <= 4 else args.partitions // 2)
    L = args.samples
    frames = L // 256 + 1
    for B in [int(x) for x in args.batches.split(",")]:
        g = torch.Generator(device="cuda").manual_seed(1)
        far = 0.1 * torch.randn(B, L, device="cuda", generator=g)
        mic = 0.5 * torch.roll(far, 37, dims=1) + 0.001 * torch.randn(B, L, device="cuda", generator=g)
        out = torch.empty_like(far)
        dbg = torch.zeros(B, nw, 12, dtype=torch.int64, device="cuda")
        cfg = A.Stage1Config(partitions=args.partitions, algo=args.algo, variant=args.variant)
        setter(None)
        A.stage1_aec(far, mic, cfg, out=out)
        torch.cuda.synchronize()
        setter(dbg.data_ptr())
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        A.stage1_aec(far, mic, cfg, out=out)
        e1.record()
        torch.cuda.synchronize()
        setter(None)
        if args.algo >= 2:      # overlap-save kernels: two phases per block, the warps swap roles every block
            names = ["R", "wait after R", "F (chain role)", "wait after F (chain)", "F (other role)", "wait after F (other)"]
            per = dbg[:, :, :6].double().sum(dim=1).mean(dim=0) / (L // 256)     # summed over the warps: cycles per block
            per[0] /= nw                                                           # R: every warp, every block
            per[1] /= nw
            per[4] /= max(nw - 1, 1)                                               # the nw - 1 warps that do not carry the chain
            per[5] /= max(nw - 1, 1)
            ms = e0.elapsed_time(e1)
            print(json.dumps({"B": B, "ms": round(ms, 3), "cycles_per_block": round(ms * 1.965e6 / (L // 256)),
                              "phases": {n: round(float(v)) for n, v in zip(names, per)}}), flush=True)
            continue
        d = dbg[:, :, :10].double()
        per_frame = d.mean(dim=0) / frames               # [nw][10] cycles per frame
        tot = per_frame.sum(dim=1)
        rec = {"B": B, "ms": e0.elapsed_time(e1), "cycles_per_frame_total": [round(float(x), 1) for x in tot]}
        for w in range(nw):
            rec[f"warp{w}"] = {NAMES[i]: round(float(per_frame[w, i]), 1) for i in range(10)}
        tt = d.sum(dim=2).sum(dim=1)                      # spread between utterances
        rec["utterance_total_min_max"] = [float(tt.min()) / nw / frames, float(tt.max()) / nw / frames]
        print(json.dumps(rec), flush=True)
        if args.placement:
            pl = dbg[:, :, 10].cpu()
            tot_c = d.sum(dim=2).cpu()                    # [B][nw] total cycles
            sm0 = int(pl[0, 0]) // 1000
            rows = [(int(pl[bb, 0]) % 1000, int(pl[bb, 1]) % 1000, bb, float(tot_c[bb, 0]) / frames)
                    for bb in range(B) if int(pl[bb, 0]) // 1000 == sm0]
            print("SM", sm0, "(hw warp slot of warp 0, of warp 1, utterance, cycles/frame):", sorted(rows), flush=True)


if __name__ == "__main__":
    main()
