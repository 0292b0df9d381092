#!/bin/bash
cd "$(dirname "$0")/.."
(time timeout 1500 python -m pytest tests -m gpu -x -q) > gpurun_out/r2_pytest_gpu_7.log 2>&1
tail -4 gpurun_out/r2_pytest_gpu_7.log
(time timeout 900 python bench.py --steps 20 --warmup 5) > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err
echo "bench rc=$?"; tail -5 gpurun_out/r2_bench_n1.err
timeout 300 python tools/sanitizer_case.py > gpurun_out/r2_sanitizer_plain.log 2>&1; echo "plain rc=$?"
timeout 1200 compute-sanitizer --tool racecheck --print-limit 20 python tools/sanitizer_case.py > gpurun_out/r2_sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?"
tail -5 gpurun_out/r2_sanitizer_racecheck.log
