#!/bin/bash
# Round-2 multi-GPU measurement batch (one gpurun --gpus 8 call): PCIe / host floor at 2, 4, 8 ranks, the host
# pipeline against it, the file pipeline (config 5 shape) and the bench line at 8 ranks.  Outputs -> gpurun_out/.
cd "$(dirname "$0")/.."
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
{ nvidia-smi topo -m; lscpu | head -25; free -g | head -2; for d in /sys/bus/pci/devices/*; do if [ -f $d/class ] && grep -q 0x0302 $d/class 2>/dev/null; then echo "$d numa=$(cat $d/numa_node) cpus=$(cat $d/local_cpulist)"; fi; done; } > $O/r2_topology_n8.txt 2>&1
timeout 300 $TR --nproc-per-node 8 --master-port 29611 tools/e2e_probe.py > $O/r2_probe_n8.jsonl 2> $O/r2_probe_n8.err; echo "probe8 rc=$?"
timeout 200 $TR --nproc-per-node 4 --master-port 29612 tools/e2e_probe.py --quick > $O/r2_probe_n4.jsonl 2> $O/r2_probe_n4.err; echo "probe4 rc=$?"
timeout 200 $TR --nproc-per-node 2 --master-port 29613 tools/e2e_probe.py --quick > $O/r2_probe_n2.jsonl 2> $O/r2_probe_n2.err; echo "probe2 rc=$?"
timeout 300 $TR --nproc-per-node 8 --master-port 29614 tools/config5_files.py --utterances 1024 > $O/r2_c5files_n8.json 2> $O/r2_c5files_n8.err; echo "c5 8 rc=$?"
timeout 300 $TR --nproc-per-node 4 --master-port 29615 tools/config5_files.py --utterances 1024 > $O/r2_c5files_n4.json 2> $O/r2_c5files_n4.err; echo "c5 4 rc=$?"
timeout 300 $TR --nproc-per-node 2 --master-port 29616 tools/config5_files.py --utterances 1024 > $O/r2_c5files_n2.json 2> $O/r2_c5files_n2.err; echo "c5 2 rc=$?"
timeout 300 $TR --nproc-per-node 8 --master-port 29617 bench.py --gpus 8 --steps 20 --warmup 3 > $O/r2_bench_n8.json 2> $O/r2_bench_n8.err; echo "bench8 rc=$?"
tail -2 $O/r2_probe_n8.jsonl | cut -c1-400
cat $O/r2_c5files_n8.json | cut -c1-600
