#!/bin/bash
# Final single-GPU pass of the round (gpurun): GPU tests, smoke, both bench arms, bench through algo 3, the file
# pipeline (config 5 as worded) with the native .ex writer and with the raw stand-in container.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2z_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2z_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2z_smoke.log 2>&1; echo "smoke rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2z_bench_ref.json 2> gpurun_out/r2z_bench_ref.err; echo "ref rc=$?"
python bench.py > gpurun_out/r2z_bench.json 2> gpurun_out/r2z_bench.err; echo "bench rc=$?"
python bench.py --algo 3 --steps 50 --no-also > gpurun_out/r2z_bench_algo3.json 2> gpurun_out/r2z_bench_algo3.err; echo "bench algo3 rc=$?"
python bench.py --impl reference --algo 3 --steps 2 --warmup 1 > gpurun_out/r2z_bench_algo3_ref.json 2>> gpurun_out/r2z_bench_algo3.err; echo "ref algo3 rc=$?"
python tools/config5_files.py --repeats 3 > gpurun_out/r2z_config5_ex.json 2> gpurun_out/r2z_config5.err; echo "config5 ex rc=$?"
python tools/config5_files.py --repeats 3 --container raw > gpurun_out/r2z_config5_raw.json 2>> gpurun_out/r2z_config5.err; echo "config5 raw rc=$?"
