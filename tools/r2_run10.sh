#!/bin/bash
cd "$(dirname "$0")/.."
out=gpurun_out/r2_run10.log
: > $out
timeout 300 python __graft_entry__.py smoke >> $out 2>&1; echo "smoke rc=$?" >> $out
timeout 300 python tools/shape_perf.py "cfg-3 E=1024" cfg-2 cfg-4 cfg-5 >> $out 2>&1; echo "rc=$?" >> $out
(time timeout 1500 python -m pytest tests -m gpu -x -q) > gpurun_out/r2_pytest_gpu_8.log 2>&1
tail -4 gpurun_out/r2_pytest_gpu_8.log >> $out
(time timeout 900 python bench.py --steps 20 --warmup 5) > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err
echo "bench rc=$?" >> $out
