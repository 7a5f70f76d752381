"""BASELINE.json configs[4]: 100 000 utterances x 10 s (16 kHz, config-2 algorithm) streamed from host memory through the
stage-1 canceller, utterance-sharded over the ranks (one process per GPU, no collective on the data path, one metrics
gather at the end).  Host buffers in, host buffers out, copies inside the timed region; the h5 writer itself is not timed
(h5py is not installed on the boxes).  A pool of 1024 distinct pinned utterances is reused as the input of every batch.

    python tools/config5_stream.py                       # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/config5_stream.py
"""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import acoustic_echo_cancellation_b200 as A  # noqa: E402
from acoustic_echo_cancellation_b200 import hostutil, sharding, synth  # noqa: E402


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        hostutil.bind_to_gpu_numa(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    total, pool, L = 100_000, 1024, 160_000
    lo, hi = sharding.shard_range(total, rank, world)
    d = synth.make_batch(1000 * rank, 8, L)                       # 8 distinct utterances tiled into the pool
    hf, hm, he = A.pinned_empty((pool, L)), A.pinned_empty((pool, L)), A.pinned_empty((pool, L))
    for i in range(pool):
        hf[i], hm[i] = d["far"][i % 8], d["mic"][i % 8]
    erle = np.empty(pool, dtype=np.float32)
    cfg = A.Stage1Config(partitions=4, algo=A.ALGO_NLMS, erle_skip_hops=125)
    pipe = A.HostPipeline(slice_utterances=128, max_samples=L, device=local)
    pipe.run(hf, hm, cfg, err=he, erle=erle)                      # warm-up
    ref = he[:8].copy()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    done, erle_sum = lo, 0.0
    while done < hi:
        nb = min(pool, hi - done)
        pipe.run(hf[:nb], hm[:nb], cfg, err=he[:nb], erle=erle[:nb])
        erle_sum += float(erle[:nb].sum())
        done += nb
    torch.cuda.synchronize()
    dt = torch.tensor([time.perf_counter() - t0], device="cuda", dtype=torch.float64)
    stats = torch.tensor([erle_sum, float(hi - lo)], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        dist.all_reduce(stats)                                   # the metrics gather of the path
    ok = bool(np.array_equal(he[:8], ref))                        # same inputs -> bit-identical outputs, batch after batch
    if rank == 0:
        secs = float(dt[0])
        print(json.dumps({"workload": "configs[4]: 100k x 10 s, 16 kHz, 4-partition FDAF-NLMS, host buffers in / out",
                          "n_gpus": world, "utterances": int(stats[1]), "seconds": secs,
                          "audio_s_per_s": total * L / 16000.0 / secs, "erle_db_mean": float(stats[0] / stats[1]),
                          "outputs_repeat_bitwise": ok,
                          "h2d_bytes": 2 * total * L * 4, "d2h_bytes": total * L * 4}), flush=True)
    pipe.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
