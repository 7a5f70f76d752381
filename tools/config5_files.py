"""BASELINE.json configs[4] as it is worded: wav files -> stage 1 -> the h5 training-set FILES, sharded over the ranks,
timed end to end (directory of 16-bit wavs in, per-utterance output files + merged tr_list.txt out), next to the
reference's loop shape (Stage2_lhm/generate_h5files/train_wav2h5.py:13-44: four decodes and one file write per
utterance, serial) on the host cores.

    python tools/config5_files.py [--utterances 2048] [--seconds 10]                      # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/config5_files.py

What is and is not measured
* The corpus is synthetic: `--pool` distinct utterances (4 wav files each, 16-bit PCM, written once, untimed) are
  hard-linked to `--utterances` ids PER RANK -- 100 000 x 10 s x 4 files would be 128 GB of wav and 384 GB of output,
  which no box here holds; the per-utterance work is the same and throughput is reported per utterance.
* Files go to --root (default /dev/shm when it has room, else the system temp dir): the page cache / tmpfs, not a
  disk array -- this measures the pipeline (decode, PCIe, kernel, container write), not a storage system.
* Container (`--container`): `auto` = h5py when it imports, else h5lite -- the package's own HDF5 writer, so the output
  IS the reference's `.ex` format (h5py / libhdf5 are absent from the image); `raw` = wav2h5.RawStore (the datasets'
  bytes + an index per utterance, same keys: the container of the round-2 tables in profiles/r2_config5_files_*);
  the JSON line says which, and one output file is read back and checked after the timed region.
* `reference_loop`: the reference's loop shape on ONE core (its scripts are single-threaded) over a bounded sample --
  scipy decode x 4 + one container write per utterance, without any filter (the reference has none) and, second
  figure, with the single-threaded C port of the stage-1 filter in the loop.
"""
import argparse
import json
import os
import shutil
import sys
import tempfile
import time
import types

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import acoustic_echo_cancellation_b200 as A  # noqa: E402
from acoustic_echo_cancellation_b200 import hostutil, ingest, wav2h5  # noqa: E402


def write_pool(folder, pool, n, sr, seed):
    from scipy.io import wavfile
    from scipy.signal import lfilter

    rng = np.random.default_rng(seed)
    os.makedirs(folder, exist_ok=True)
    for u in range(pool):
        far = lfilter([1.0], [1.0, -0.9], rng.standard_normal(n))
        far = 0.5 * far / np.abs(far).max()
        h = rng.standard_normal(512) * np.exp(-np.arange(512) / 80.0)
        h *= 0.5 / np.linalg.norm(h)
        echo = np.convolve(far, h)[:n]
        near = 0.05 * rng.standard_normal(n)
        sig = {"farend_speech": far, "echo": echo, "nearend_speech": near, "nearend_mic": echo + near}
        for k, v in sig.items():
            pcm = np.clip(np.rint(v * 32768.0), -32768, 32767).astype(np.int16)
            wavfile.write(os.path.join(folder, wav2h5.WAV_PATTERNS[k].format(idx=f"p{u}")), sr, pcm)


def link_ids(pool_dir, train_dir, pool, first, count):
    os.makedirs(train_dir, exist_ok=True)
    for i in range(first, first + count):
        for k, pat in wav2h5.WAV_PATTERNS.items():
            os.link(os.path.join(pool_dir, pat.format(idx=f"p{i % pool}")), os.path.join(train_dir, pat.format(idx=str(i))))


def reference_loop(train_dir, out_dir, ids, sr, store, with_filter):
    """train_wav2h5.py:13-44 in shape: serial, four decodes and one file per utterance."""
    from oracle import aec_oracle as O       # bench-only use of the oracle (CPU baseline leg)
    from oracle import c_oracle as CO

    os.makedirs(out_dir, exist_ok=True)
    t0 = time.perf_counter()
    for idx in ids:
        sig = {k: ingest.load_wav(os.path.join(train_dir, wav2h5.WAV_PATTERNS[k].format(idx=idx)), sr) for k in wav2h5.KEYS}
        w = store.File(os.path.join(out_dir, "tr_" + idx + ".ex"), "w")
        for k in wav2h5.KEYS:
            w.create_dataset(k, data=sig[k].astype(np.float32), shape=sig[k].shape, chunks=True)
        if with_filter:
            r = CO.stage1(sig["farend_speech"][None], sig["nearend_mic"][None], O.AecConfig(), n_threads=1)
            w.create_dataset("stage1_error", data=r["err"][0], shape=r["err"][0].shape, chunks=True)
            w.create_dataset("stage1_echo", data=r["echo"][0], shape=r["echo"][0].shape, chunks=True)
        w.close()
    return (time.perf_counter() - t0) / len(ids)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--utterances", type=int, default=2048, help="per rank")
    ap.add_argument("--pool", type=int, default=128)
    ap.add_argument("--seconds", type=float, default=10.0)
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--decode-threads", type=int, default=0, help="0 = cores available to this rank")
    ap.add_argument("--write-threads", type=int, default=0)
    ap.add_argument("--root", default=None)
    ap.add_argument("--container", default="auto", choices=["auto", "raw"])
    ap.add_argument("--repeats", type=int, default=3, help="timed conversions of the set; the median is the headline (phase seconds: last run)")
    ap.add_argument("--ref-sample", type=int, default=48)
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        hostutil.bind_to_gpu_numa(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    cores = len(os.sched_getaffinity(0))
    per_rank_cores = max(2, cores // max(1, world)) if world > 1 else cores
    dthreads = args.decode_threads or per_rank_cores
    wthreads = args.write_threads or per_rank_cores
    sr, n = 16000, int(args.seconds * 16000)
    total = args.utterances * world
    need = total * n * 4 * 6 + args.pool * n * 2 * 4
    base = args.root
    if base is None:
        base = "/dev/shm" if os.path.isdir("/dev/shm") and shutil.disk_usage("/dev/shm").free > 1.3 * need else tempfile.gettempdir()
    work = os.path.join(base, "aec_config5_files")
    pool_dir, train_dir = os.path.join(work, "pool"), os.path.join(work, "train")
    h5_dir, list_dir = os.path.join(work, "h5"), os.path.join(work, "lists")
    if rank == 0:
        shutil.rmtree(work, ignore_errors=True)
        os.makedirs(list_dir)
        write_pool(pool_dir, args.pool, n, sr, 5)
        link_ids(pool_dir, train_dir, args.pool, 0, total)
    if world > 1:
        dist.barrier()
    if args.container == "raw":
        store, container = wav2h5.RawStore(), "RawStore (raw stand-in container)"
    else:
        try:
            import h5py  # type: ignore
            store, container = h5py, "h5py"
        except ImportError:
            from acoustic_echo_cancellation_b200 import h5lite
            store, container = h5lite, "h5lite (HDF5 written by the package itself: h5py is not installed)"
    ns = types.SimpleNamespace(train_path=train_dir, h5_path=h5_dir, list_path=list_dir, sr=sr)
    runner = wav2h5.default_runner(slice_utterances=128, device=local)
    # warm-up: CUDA context, page-locked buffers, kernel load (a small separate run into a scratch folder)
    warm = types.SimpleNamespace(train_path=os.path.join(work, f"warm{rank}"), h5_path=os.path.join(work, f"warm_h5_{rank}"),
                                 list_path=os.path.join(work, f"warm_l_{rank}"), sr=sr)
    os.makedirs(warm.list_path, exist_ok=True)
    # (every rank's shard of them is three full batches: all three decoder buffer sets and both output sets of the
    #  runner get their final size -- page-locked allocations are slow and would otherwise land in the timed run)
    link_ids(pool_dir, warm.train_path, args.pool, 10 ** 9 + rank * 100000, 3 * args.batch * world)
    wav2h5.create_h5_train(warm, runner=runner, batch=args.batch, h5=store, decode_threads=dthreads,
                           write_threads=wthreads, pinned=True)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    runs = []
    for rep in range(max(1, args.repeats)):                  # the whole conversion, `--repeats` times; the MEDIAN is the
        if rep:                                               # headline, every wall time is in `wall_s_runs`
            if world > 1:
                dist.barrier()
            if rank == 0:
                shutil.rmtree(h5_dir, ignore_errors=True)
            if world > 1:
                dist.barrier()
        st = {}
        t0 = time.perf_counter()
        merged = wav2h5.create_h5_train(ns, runner=runner, batch=args.batch, h5=store, decode_threads=dthreads,
                                        write_threads=wthreads, pinned=True, stats=st)
        dt = torch.tensor([time.perf_counter() - t0], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        runs.append(float(dt[0]))
    secs = sorted(runs)[len(runs) // 2]                      # headline = the median run
    ok = len(merged) == total and all(os.path.exists(p) for p in merged[:: max(1, total // 64)])
    if ok and rank == 0 and args.container != "raw" and container.startswith("h5lite"):
        from acoustic_echo_cancellation_b200 import h5lite      # one output file read back: keys, shapes, finite values
        with h5lite.File(merged[len(merged) // 2], "r") as r:
            ok = (set(r) == set(wav2h5.KEYS) | {"stage1_error", "stage1_echo"} and r["stage1_error"].shape == (n,)
                  and bool(np.isfinite(r["stage1_error"][:]).all()) and r["nearend_mic"].dtype == np.float32)
    if rank == 0:
        line = {"workload": "configs[4] shape: wav dir -> stage 1 (4-partition FDAF-NLMS) -> one output file per utterance + "
                            "tr_list.txt, sharded over the ranks", "n_gpus": world, "utterances": total,
                "seconds_per_utterance_audio": args.seconds, "wall_s": secs, "wall_s_runs": runs, "utterances_per_s": total / secs,
                "audio_s_per_s": total * args.seconds / secs, "container": container, "files_root": base,
                "decode_threads_per_rank": dthreads, "write_threads_per_rank": wthreads, "host_cores": cores,
                "rank0_phase_seconds": {k: round(v, 3) for k, v in st.items() if k.endswith("_s")},
                "rank0_pcm16_batches": st.get("pcm16_batches"), "all_files_listed_and_present": bool(ok),
                "bytes_in_per_utterance": 4 * n * 2, "bytes_out_per_utterance": 6 * n * 4}
        # the reference's loop shape, one core, bounded sample
        ids = [str(i) for i in range(min(args.ref_sample, total))]
        plain = reference_loop(train_dir, os.path.join(work, "ref_plain"), ids, sr, store, with_filter=False)
        filt = reference_loop(train_dir, os.path.join(work, "ref_filter"), ids[: max(8, len(ids) // 4)], sr, store,
                              with_filter=True)
        line["reference_loop"] = {
            "sample": len(ids), "threads": 1,
            "utterances_per_s_no_filter": 1.0 / plain, "audio_s_per_s_no_filter": args.seconds / plain,
            "utterances_per_s_with_c_port_filter": 1.0 / filt, "audio_s_per_s_with_c_port_filter": args.seconds / filt,
            "note": "serial loop of train_wav2h5.py:13-44 (scipy decode; the reference uses librosa, absent here); the "
                    "reference has no filter -- the second figure adds the single-threaded C port"}
        line["speedup_vs_reference_loop_with_filter"] = line["utterances_per_s"] * filt
        print(json.dumps(line), flush=True)
        shutil.rmtree(work, ignore_errors=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
