#!/usr/bin/env python
"""bench.py -- audio-seconds per second of the stage-1 STFT + FDAF echo canceller on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config 2|3]
    (N > 1: launched by torchrun, one rank per GPU)

One "step" = one pass of the hot path over one batch of synthetic utterances:
  workload  BASELINE.json configs[1]: 1024 synthetic 10 s utterances per GPU, 16 kHz, frame 512 /
            hop 256, 4-partition FDAF-NLMS  (weak scaling: every rank owns its own 1024).
  value     whole-job audio-seconds processed per second, inputs resident in HBM, device-timed
            (CUDA events on the launching stream, max over ranks).
  e2e       same metric through the host-buffer C-ABI call (aec_stage1_run_host): pinned host
            inputs -> H2D -> kernel -> D2H of the error signal, all inside the timed region.
  roofline  the fused stage-1 kernel against the MEASURED FP32 FFMA peak of this GPU (binding
            roofline) with the HBM fraction (MEASURED_PEAKS.json) beside it.
  cpu_baseline  the C oracle (builder-authored port; the reference has no stage-1 filter) on the
            box's host cores over a bounded sample of the same workload.
`--impl reference` times that CPU port alone (there is no reference implementation of this path).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    2: dict(name="configs[1]: 1024 x 10 s utterances/GPU, 16 kHz, frame 512 hop 256, 4-partition FDAF-NLMS",
            B=1024, L=160000, P=4, algo=0, frame=512, sr=16000),
    3: dict(name="configs[2]: 4096 x 10 s utterances/GPU, 16 kHz, frame 512 hop 256, 16-partition Kalman FDAF",
            B=4096, L=160000, P=16, algo=1, frame=512, sr=16000),
    4: dict(name="configs[3]: 1024 x 10 s utterances/GPU, 48 kHz, frame 1024 hop 512, 8-partition FDAF-NLMS",
            B=1024, L=480000, P=8, algo=0, frame=1024, sr=48000),
}


def flops_per_frame(N, P, algo):
    import math
    K = N // 2 + 1
    common = 7.5 * N * math.log2(N) + 5 * N
    return common + (16 * K * P + 12 * K if algo == 0 else 31 * K * P + 11 * K)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 20 ms.  Samples are time-stamped on
    arrival; `summary(t0, t1)` reports the ones that fell inside the timed region [t0, t1] and, when
    the region is too short for three samples, the whole under-load window (the sampler keeps running
    over identical untimed launches after the timed region) -- and says which."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self, t0, t1, load0, load1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

        def collect(a, b):
            sm, mx, power, reasons = [], [], [], set()
            for ts, r in self.rows:
                if not (a <= ts <= b):
                    continue
                f = [x.strip() for x in r.split(",")]
                if len(f) < 7:
                    continue
                try:
                    sm.append(float(f[0])); mx.append(float(f[1])); power.append(float(f[2]))
                except ValueError:
                    continue
                for nm, v in zip(names, f[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            return sm, mx, power, reasons

        sm, mx, power, reasons = collect(t0, t1)
        window = "timed region"
        if len(sm) < 3:
            sm, mx, power, reasons = collect(load0, load1)
            window = "timed region + identical untimed launches around it (region shorter than 3 samples)"
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "power_w_max": max(power),
                "samples": len(sm), "window": window, "reasons": sorted(reasons)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "MEASURED_PEAKS.json (driver-measured copy bandwidth)"
        except Exception:
            pass
    return 6650.0, "fallback of B200_PROFILING.md (MEASURED_PEAKS.json absent)"


def make_inputs_device(torch, B, L, seed, P, SR=16000, hop=256):
    """SURVEY.md 8d recipe on the device (seeded): speech-like far end, exponentially decaying random
    RIR of P*256 taps, mic = echo + noise at -40 dB.  Untimed set-up."""
    g = torch.Generator(device="cuda").manual_seed(seed)
    far = torch.empty(B, L, device="cuda")
    mic = torch.empty(B, L, device="cuda")
    t = torch.arange(L, device="cuda", dtype=torch.float32) / SR
    env = 0.5 - 0.5 * torch.cos(2 * torch.pi * 4.0 * t)
    nfft = 1 << (L + P * hop).bit_length()
    lp = torch.fft.rfft(0.9 ** torch.arange(128, device="cuda", dtype=torch.float32), n=nfft)
    tau = P * hop / 6.9
    dec = torch.exp(-torch.arange(P * hop, device="cuda", dtype=torch.float32) / tau)
    step = 64 if L <= 200000 else 16
    for b0 in range(0, B, step):
        nb = min(step, B - b0)
        x = torch.randn(nb, L, device="cuda", generator=g)
        x = torch.fft.irfft(torch.fft.rfft(x, n=nfft) * lp, n=nfft)[:, :L] * env
        x = 0.5 * x / x.abs().amax(dim=1, keepdim=True)
        h = torch.randn(nb, P * hop, device="cuda", generator=g) * dec
        h = 0.5 * h / h.norm(dim=1, keepdim=True)
        echo = torch.fft.irfft(torch.fft.rfft(x, n=nfft) * torch.fft.rfft(h, n=nfft), n=nfft)[:, :L]
        noise = torch.randn(nb, L, device="cuda", generator=g) * echo.pow(2).mean(dim=1, keepdim=True).sqrt() * 0.01
        far[b0:b0 + nb] = x
        mic[b0:b0 + nb] = echo + noise
    return far, mic


def make_inputs_host(np, B, L, seed):
    """Host-only inputs for the CPU arm (no GPU needed): same shape/scale, numpy PCG64."""
    rng = np.random.default_rng(seed)
    far = (0.15 * rng.standard_normal((B, L))).astype(np.float32)
    h = (rng.standard_normal(64) * np.exp(-np.arange(64) / 10.0)).astype(np.float32)
    h *= 0.5 / np.linalg.norm(h)
    mic = np.empty_like(far)
    for b in range(B):
        mic[b] = np.convolve(far[b], h)[:L]
    mic += (0.001 * rng.standard_normal((B, L))).astype(np.float32)
    return far, mic


def cpu_arm(np, far, mic, wl, steps, warmup):
    """C oracle (port) on all host threads; returns (audio_s_per_s, threads, ms_per_step)."""
    from oracle import aec_oracle as O
    from oracle import c_oracle as CO

    cfg = O.AecConfig(frame=wl["frame"], partitions=wl["P"], algo=wl["algo"], delta=1e-6 * wl["frame"])
    # every core this process may run on (torchrun exports OMP_NUM_THREADS=1; the explicit thread
    # count passed to the C entry overrides it)
    try:
        threads = len(os.sched_getaffinity(0))
    except AttributeError:
        threads = os.cpu_count() or 1
    out = (np.zeros_like(far), None, np.zeros(far.shape[0], dtype=np.float32))   # pre-faulted outputs
    for _ in range(warmup):
        CO.stage1(far, mic, cfg, want_echo=False, out=out, n_threads=threads)
    t0 = time.perf_counter()
    for _ in range(steps):
        CO.stage1(far, mic, cfg, want_echo=False, out=out, n_threads=threads)
    dt = (time.perf_counter() - t0) / max(steps, 1)
    return far.shape[0] * far.shape[1] / wl["sr"] / dt, threads, dt * 1e3


def run_reference(args, wl):
    """`--impl reference`: the CPU arm.  The reference repository has no implementation of this
    path (no FDAF at all), so the arm is the builder-authored C port on the host cores."""
    import numpy as np

    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    sample = int(min(512, max(32, 16 * cores)))
    far, mic = make_inputs_host(np, sample, wl["L"], 99)
    v, threads, ms = cpu_arm(np, far, mic, wl, args.steps, max(args.warmup, 1))
    line = {
        "impl": "reference", "metric": "audio_seconds_per_second_stage1_aec", "value": v, "unit": "audio-s/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl["name"], "sample": f"{sample} utterances x 10 s per step"},
        "cpu_baseline": {"value": v, "unit": "audio-s/s", "cores": threads, "kind": "port",
                         "sample": f"{sample} x 10 s utterances per step, C oracle (oracle/csrc/aec_oracle.c), "
                                   f"OpenMP over utterances; the reference has no CPU FDAF to time"},
        "e2e": {"value": v, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=sorted(WORKLOADS))
    ap.add_argument("--variant", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-numa-bind", action="store_true")
    args = ap.parse_args()
    wl = WORKLOADS[args.config]
    if args.impl == "reference":
        return run_reference(args, wl)
    args.warmup = max(args.warmup, 3)

    import numpy as np
    import torch
    import torch.distributed as dist

    import acoustic_echo_cancellation_b200 as A
    from acoustic_echo_cancellation_b200 import sharding

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a GPU: acoustic_echo_cancellation_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    numa_cpus = None
    if world > 1 and not args.no_numa_bind:
        from acoustic_echo_cancellation_b200 import hostutil
        numa_cpus = hostutil.bind_to_gpu_numa(local)     # host buffers of the e2e leg land GPU-local
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    B, L, P, algo = wl["B"], wl["L"], wl["P"], wl["algo"]
    SR, FRAME = wl["sr"], wl["frame"]
    HOP = FRAME // 2
    far, mic = make_inputs_device(torch, B, L, 1000 + rank, P, SR, HOP)
    err = torch.empty_like(far)
    cfg = A.Stage1Config(frame=FRAME, partitions=P, algo=algo, erle_skip_hops=125, variant=args.variant)
    n_total = B * world

    def step():
        _, erle = A.stage1_aec(far, mic, cfg, out=err, return_erle=True)
        if world > 1:   # the path's only collective: per-utterance metrics, never the signals
            erle = sharding.gather_metrics(erle, n_total)
        return erle

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    fp32_peak = A.fp32_peak_tflops()
    barrier()

    # ---- timed region: K steps, device-timed, inputs (1.3 GB/GPU) far larger than the 126 MB L2 ----
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.1)
    load0 = time.perf_counter()
    for _ in range(max(3, int(0.15 / 0.002))):      # ~0.15 s of identical launches so the sampler sees load
        A.stage1_aec(far, mic, cfg, out=err, return_erle=True)
    barrier()
    A.launch_count(reset=True)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    wall0 = time.perf_counter()
    ev0.record()
    pending = []
    for i in range(args.steps):
        kev[i][0].record()
        erle = A.stage1_aec(far, mic, cfg, out=err, return_erle=True)[1]
        kev[i][1].record()
        if world > 1:     # metrics gather of step i runs on NCCL's stream under the kernel of step i + 1
            pending.append(sharding.gather_metrics(erle, n_total, async_op=True))
    for erle, work in pending:
        if work is not None:
            work.wait()   # every step's gathered metrics are complete inside the timed region
    ev1.record()
    barrier()
    wall1 = time.perf_counter()
    wall = wall1 - wall0
    launches = A.launch_count()
    clocks = None
    if rank == 0:
        if wall < 0.3:                               # keep the GPU under the same load while sampling
            t_end = time.perf_counter() + 0.3
            while time.perf_counter() < t_end:
                A.stage1_aec(far, mic, cfg, out=err, return_erle=True)
                torch.cuda.synchronize()
        load1 = time.perf_counter()
        time.sleep(0.05)
        sampler.stop()
        clocks = sampler.summary(wall0, wall1, load0, load1)
    barrier()
    total_ms = ev0.elapsed_time(ev1)
    kern_ms = sum(a.elapsed_time(b) for a, b in kev) / args.steps
    tmax = torch.tensor([total_ms, kern_ms], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    total_ms, kern_ms = float(tmax[0]), float(tmax[1])
    ms_per_step = total_ms / args.steps
    audio_s_step = n_total * L / SR
    value = audio_s_step / (ms_per_step * 1e-3)
    erle_mean = float(erle.float().mean())
    finite = bool(torch.isfinite(err).all())

    # ---- e2e: pinned host buffers through aec_stage1_run_host, copies inside the timed region ----
    e2e = None
    if not args.no_e2e:
        hf, hm, he = A.pinned_empty((B, L)), A.pinned_empty((B, L)), A.pinned_empty((B, L))
        herle = np.empty(B, dtype=np.float32)
        hf[:] = far.cpu().numpy()
        hm[:] = mic.cpu().numpy()
        pipe = A.HostPipeline(slice_utterances=min(128, B), max_samples=L, device=local)
        for _ in range(2):
            pipe.run(hf, hm, cfg, err=he, erle=herle)
        barrier()
        k_e2e = max(3, min(args.steps, 10))
        t0 = time.perf_counter()
        for _ in range(k_e2e):
            pipe.run(hf, hm, cfg, err=he, erle=herle)     # returns when the outputs are in host memory
        barrier()
        dt = torch.tensor([(time.perf_counter() - t0) / k_e2e], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        e2e = {"value": audio_s_step / float(dt[0]), "unit": "audio-s/s",
               "h2d_bytes_per_step": 2 * B * L * 4, "d2h_bytes_per_step": B * L * 4 + B * 4,
               "ms_per_step": float(dt[0]) * 1e3, "steps": k_e2e,
               "api": "aec_stage1_run_host (HostPipeline.run), float32 pinned host memory, 128-utterance slices (tapered at "
                      "the end), 4 slices in flight; PCIe-bound (H2D %.2f GB/step at ~50 GB/s with D2H running)" % (2 * B * L * 4 / 1e9)}
        e2e_match = bool(np.array_equal(he, err.cpu().numpy()))
        # wav-ingest variant: 16-bit PCM host buffers (what the wav files hold), converted on the GPU
        h16f = A.pinned_empty((B, L), dtype=np.int16)
        h16m = A.pinned_empty((B, L), dtype=np.int16)
        h16f[:] = np.clip(np.rint(hf * 32768.0), -32768, 32767).astype(np.int16)
        h16m[:] = np.clip(np.rint(hm * 32768.0), -32768, 32767).astype(np.int16)
        for _ in range(2):
            pipe.run(h16f, h16m, cfg, err=he, erle=herle)
        barrier()
        t0 = time.perf_counter()
        for _ in range(k_e2e):
            pipe.run(h16f, h16m, cfg, err=he, erle=herle)
        barrier()
        dt16 = torch.tensor([(time.perf_counter() - t0) / k_e2e], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(dt16, op=dist.ReduceOp.MAX)
        e2e["pcm16_variant"] = {"value": audio_s_step / float(dt16[0]), "unit": "audio-s/s",
                                "h2d_bytes_per_step": 2 * B * L * 2, "d2h_bytes_per_step": B * L * 4 + B * 4,
                                "ms_per_step": float(dt16[0]) * 1e3,
                                "api": "aec_stage1_run_host_pcm16: int16 PCM in (wav sample format), float32 out"}
        pipe.close()
    else:
        hf = hm = None
        e2e_match = None

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant (only) kernel ----
    frames = L // HOP + 1
    flops_launch = flops_per_frame(FRAME, P, algo) * frames * B
    bytes_launch = 3 * 4 * L * B
    hbm_peak, hbm_src = measured_peaks()
    tf = flops_launch / (kern_ms * 1e-3) / 1e12
    gbs = bytes_launch / (kern_ms * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get(f"config{args.config}")
        except Exception:
            traffic = None
    roofline = {"bound": "fp32", "achieved": tf, "peak": fp32_peak, "unit": "TFLOP/s", "frac": tf / fp32_peak,
                "traffic": traffic, "kernel": "aec::stage1_n%d_kernel" % FRAME, "kernel_ms": kern_ms,
                "flops_per_launch": flops_launch, "bytes_per_launch": bytes_launch,
                "peak_source": "FFMA probe measured live in this run (aec_bench_fp32_peak); "
                               "MEASURED_PEAKS.json carries no FP32 figure",
                "hbm": {"achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak,
                        "peak_source": hbm_src}}

    if args.config == 2:
        # context for `frac` (static analysis of the measured SASS, profiles/r1_fp32_issue_rates.txt): FFMAs with three
        # distinct register sources issue at 0.64 per cycle per scheduler on this part; with them counted at that rate the
        # instruction stream of this kernel cannot exceed ~9.1 M audio-s/s per GPU (0.44 of the FFMA peak)
        ceiling = 9.1e6
        roofline["issue_limited_ceiling"] = {"audio_s_per_s_per_gpu": ceiling, "frac_of_ceiling": (value / world) / ceiling,
                                             "source": "profiles/r1_fp32_issue_rates.txt, DESIGN.md section 5 (Roofline)"}

    # ---- CPU baseline: the C port on this box's cores, bounded sample of the same workload ----
    cpu = None
    if not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        sample = int(min(512, max(32, 16 * cores), B))
        if hf is not None:
            sf, sm_ = np.array(hf[:sample]), np.array(hm[:sample])
        else:
            sf, sm_ = far[:sample].cpu().numpy(), mic[:sample].cpu().numpy()
        v, threads, _ = cpu_arm(np, sf, sm_, wl, 2, 1)
        cpu = {"value": v, "unit": "audio-s/s", "cores": threads, "kind": "port",
               "sample": f"first {sample} utterances x 10 s of the same workload, 2 timed passes, C oracle "
                         f"(builder-authored port: the reference has no stage-1 filter), OpenMP over utterances"}

    line = {
        "metric": "audio_seconds_per_second_stage1_aec", "value": value, "unit": "audio-s/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic (seeded on device: speech-like far end, random decaying RIR, -40 dB noise)",
        "config": {"workload": wl["name"], "utterances_per_gpu": B, "samples": L, "sample_rate": SR,
                   "frame": FRAME, "hop": HOP, "partitions": P, "algo": "nlms" if algo == 0 else "kalman",
                   "l2": "inputs %.1f GB/GPU per step >> 126 MB L2 (no flush needed)" % (2 * B * L * 4 / 1e9),
                   "parallelism": f"utterance-sharded x{world}, metrics-only all_gather",
                   "host_cpus_rank0": (f"{len(numa_cpus)} GPU-local CPUs" if numa_cpus else "unbound")},
        "per_gpu": value / world,
        "e2e": e2e, "e2e_bitwise_equal_to_device_path": e2e_match,
        "gpu_launches": int(launches),
        "roofline": roofline,
        "roofline_hbm": {"bound": "hbm", "achieved": roofline["hbm"]["achieved"], "peak": roofline["hbm"]["peak"],
                         "unit": "GB/s", "frac": roofline["hbm"]["frac"], "traffic": traffic,
                         "note": "secondary: the binding roofline of this path is FP32 (see roofline)"},
        "cpu_baseline": cpu, "clocks": clocks,
        "wall_s_timed_region": wall, "erle_db_mean": erle_mean, "outputs_finite": finite,
        "parity": "FDAF recurrence: parity UNPINNED (no reference implementation); STFT/iSTFT pinned by golden vectors",
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
