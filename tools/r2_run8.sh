#!/bin/bash
cd "$(dirname "$0")/.."
(time timeout 1200 python -m pytest tests -m gpu -x -q) > gpurun_out/r2_pytest_gpu_6.log 2>&1
tail -4 gpurun_out/r2_pytest_gpu_6.log
(time timeout 600 python bench.py --steps 10 --warmup 3) > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err
echo "bench rc=$?"; tail -5 gpurun_out/r2_bench_n1.err
timeout 300 python tools/shape_perf.py "cfg-3 E=1024" cfg-2 cfg-4 cfg-5 > gpurun_out/r2_shape_perf_2.log 2>&1
