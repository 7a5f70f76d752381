#!/bin/bash
# Round-2 second multi-GPU batch: the file pipeline (config 5 shape) at 1, 2, 4, 8 ranks with warmed buffers.
cd "$(dirname "$0")/.."
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 200 python tools/config5_files.py --utterances 2048 > $O/r2b_c5files_n1.json 2> $O/r2b_c5files_n1.err; echo "c5 1 rc=$?"
timeout 300 $TR --nproc-per-node 2 --master-port 29621 tools/config5_files.py --utterances 2048 > $O/r2b_c5files_n2.json 2> $O/r2b_c5files_n2.err; echo "c5 2 rc=$?"
timeout 300 $TR --nproc-per-node 4 --master-port 29622 tools/config5_files.py --utterances 2048 > $O/r2b_c5files_n4.json 2> $O/r2b_c5files_n4.err; echo "c5 4 rc=$?"
timeout 300 $TR --nproc-per-node 8 --master-port 29623 tools/config5_files.py --utterances 2048 > $O/r2b_c5files_n8.json 2> $O/r2b_c5files_n8.err; echo "c5 8 rc=$?"
for n in 1 2 4 8; do grep -o '"utterances_per_s": [0-9.]*' $O/r2b_c5files_n$n.json; done
