"""PCIe / host-memory floor of the host-buffer path, at 1..8 ranks, and the host pipeline against it.

    python tools/e2e_probe.py [--quick]                                                       # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/e2e_probe.py

Part 1 (raw floor): every active rank copies a config-2-sized signal set (1024 x 10 s float32 = 655 MB per
signal) between page-locked host memory and its GPU with plain cudaMemcpyAsync -- H2D only, D2H only, both
directions at once, and the mix the canceller needs (2 signals up, 1 down; 1 int16-sized signal pair up, 1 down)
-- for default and write-combined host buffers, and for subsets of the ranks (which GPUs share a host path).
All ranks start together (barrier) and the slowest rank's time counts.
Part 2: HostPipeline.run (aec_stage1_run_host / _pcm16) over slots x slice x ramp, same timing.
One JSON object per line on rank 0's stdout; profiles/r2_pcie_floor.md is written from them.
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import acoustic_echo_cancellation_b200 as A  # noqa: E402
from acoustic_echo_cancellation_b200 import hostutil  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--no-pipeline", action="store_true")
    ap.add_argument("--no-subsets", action="store_true")
    ap.add_argument("--numa-bind", action="store_true")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if args.numa_bind:
        hostutil.bind_to_gpu_numa(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def emit(obj):
        if rank == 0:
            obj["n_ranks"] = world
            print(json.dumps(obj), flush=True)

    def timed(fn, active=True, reps=3, finish=None):
        """max over ranks of the per-repetition time of fn (inactive ranks only take part in the barriers)"""
        if active:
            fn()
            if finish:
                finish()
        barrier()
        t0 = time.perf_counter()
        if active:
            for _ in range(reps):
                fn()
            if finish:
                finish()
        torch.cuda.synchronize()
        dt = torch.tensor([(time.perf_counter() - t0) / reps if active else 0.0], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        return float(dt[0])

    B, L = 1024, 160000
    sig = B * L * 4 / 1e9                        # GB per float32 signal set
    dev = [torch.empty(B, L, device="cuda") for _ in range(3)]
    s_up, s_up2, s_dn = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()

    def host_set(flags):
        return [torch.from_numpy(A.pinned_empty((B, L), np.float32, flags)) for _ in range(2)]

    h_out = torch.from_numpy(A.pinned_empty((B, L), np.float32))
    for kind, flags in (("default", 0), ("write_combined", A.HOST_WRITE_COMBINED)):
        h_in = host_set(flags)
        for t in h_in:
            t.fill_(0.25)

        def up1():
            with torch.cuda.stream(s_up):
                dev[0].copy_(h_in[0], non_blocking=True)

        def dn1():
            with torch.cuda.stream(s_dn):
                h_out.copy_(dev[2], non_blocking=True)

        def both():
            up1()
            dn1()

        def mix_f32():          # what one float32 step moves: two signals up, one down
            with torch.cuda.stream(s_up):
                dev[0].copy_(h_in[0], non_blocking=True)
                dev[1].copy_(h_in[1], non_blocking=True)
            dn1()

        def mix_pcm16():        # int16 inputs: half of each signal up, one float32 signal down
            with torch.cuda.stream(s_up):
                dev[0][:B // 2].copy_(h_in[0][:B // 2], non_blocking=True)
                dev[1][:B // 2].copy_(h_in[1][:B // 2], non_blocking=True)
            dn1()

        subsets = [list(range(world))]
        if world == 8 and kind == "default" and not args.no_subsets and not args.quick:
            subsets += [[0], [0, 1], [0, 2], [0, 4], [0, 1, 2, 3], [4, 5, 6, 7], [0, 2, 4, 6]]
        elif world == 4 and kind == "default" and not args.no_subsets and not args.quick:
            subsets += [[0], [0, 1], [0, 2], [2, 3]]
        for ranks in subsets:
            act = rank in ranks
            n = len(ranks)
            row = {"probe": "raw", "host_memory": kind, "ranks": ranks}
            dt = timed(up1, act)
            row["h2d_gbs_total"] = n * sig / dt
            dt = timed(dn1, act)
            row["d2h_gbs_total"] = n * sig / dt
            dt = timed(both, act)
            row["h2d_and_d2h_gbs_each_total"] = n * sig / dt
            dt = timed(mix_f32, act)
            row["step_f32_ms"] = dt * 1e3
            row["step_f32_gbs_total"] = n * 3 * sig / dt
            row["floor_f32_audio_s_per_s_total"] = n * B * 10.0 / dt
            dt = timed(mix_pcm16, act)
            row["step_pcm16_ms"] = dt * 1e3
            row["step_pcm16_gbs_total"] = n * 2 * sig / dt
            row["floor_pcm16_audio_s_per_s_total"] = n * B * 10.0 / dt
            emit(row)
        del h_in
    if args.no_pipeline:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- part 2: the host pipeline ----
    del dev
    rng = np.random.default_rng(rank)
    hf, hm, he = A.pinned_empty((B, L)), A.pinned_empty((B, L)), A.pinned_empty((B, L))
    hf[:] = rng.standard_normal((B, L)).astype(np.float32) * 0.1
    hm[:] = hf * 0.5
    h16f, h16m = A.pinned_empty((B, L), np.int16), A.pinned_empty((B, L), np.int16)
    h16f[:] = np.clip(np.rint(hf * 32768.0), -32768, 32767).astype(np.int16)
    h16m[:] = np.clip(np.rint(hm * 32768.0), -32768, 32767).astype(np.int16)
    wf = A.pinned_empty((B, L), np.int16, A.HOST_WRITE_COMBINED)
    wm = A.pinned_empty((B, L), np.int16, A.HOST_WRITE_COMBINED)
    wf[:] = h16f
    wm[:] = h16m
    cfg = A.Stage1Config()
    grid = [(4, 128, True), (4, 128, False), (2, 128, True), (6, 128, True), (8, 64, True), (4, 64, True),
            (4, 256, True), (2, 256, True), (3, 512, True)]
    if args.quick:
        grid = [(4, 128, True), (4, 128, False), (2, 256, True)]
    for slots, sl, ramp in grid:
        pipe = A.HostPipeline(sl, L, device=local, slots=slots, ramp=ramp)
        row = {"probe": "pipeline", "slots": slots, "slice": sl, "ramp": ramp}
        dt = timed(lambda: pipe.run(hf, hm, cfg, err=he))
        row["f32_ms"], row["f32_audio_s_per_s_total"] = dt * 1e3, world * B * 10.0 / dt
        dt = timed(lambda: pipe.run(h16f, h16m, cfg, err=he))
        row["pcm16_ms"], row["pcm16_audio_s_per_s_total"] = dt * 1e3, world * B * 10.0 / dt
        dt = timed(lambda: pipe.run(wf, wm, cfg, err=he))
        row["pcm16_wc_ms"], row["pcm16_wc_audio_s_per_s_total"] = dt * 1e3, world * B * 10.0 / dt
        # streaming: steps issued back to back in deferred mode, one wait at the end (the tail of a step under the next ramp)
        dt = timed(lambda: pipe.run(h16f, h16m, cfg, err=he, wait=False), reps=4, finish=pipe.wait)
        row["pcm16_streaming_ms"], row["pcm16_streaming_audio_s_per_s_total"] = dt * 1e3, world * B * 10.0 / dt
        dt = timed(lambda: pipe.run(hf, hm, cfg, err=he, wait=False), reps=4, finish=pipe.wait)
        row["f32_streaming_ms"], row["f32_streaming_audio_s_per_s_total"] = dt * 1e3, world * B * 10.0 / dt
        emit(row)
        pipe.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
