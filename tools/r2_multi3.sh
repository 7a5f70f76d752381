#!/bin/bash
# Round-2 third multi-GPU batch: the final bench line at 8 ranks (+ reference arm) and the pipeline rows of the probe
# (per-call and streaming) at 8 ranks.
cd "$(dirname "$0")/.."
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 $TR --nproc-per-node 8 --master-port 29631 bench.py --gpus 8 --steps 50 --warmup 3 > $O/r2c_bench_n8.json 2> $O/r2c_bench_n8.err; echo "bench8 rc=$?"
timeout 200 $TR --nproc-per-node 8 --master-port 29632 bench.py --impl reference --gpus 8 --steps 3 --warmup 1 > $O/r2c_bench_n8_ref.json 2> $O/r2c_bench_n8_ref.err; echo "ref8 rc=$?"
timeout 300 $TR --nproc-per-node 8 --master-port 29633 tools/e2e_probe.py --quick --no-subsets > $O/r2c_probe_n8.jsonl 2> $O/r2c_probe_n8.err; echo "probe8 rc=$?"
grep pipeline $O/r2c_probe_n8.jsonl | cut -c1-700
