#!/bin/bash
cd "$(dirname "$0")/.."
P=embodied-one-shot-video-recognition_b200
timeout 1500 python tools/ab_perf.py $P/libeosvr_prev.so $P/libeosvr.so 2 > gpurun_out/r2_ab.log 2>&1
cat gpurun_out/r2_ab.log
(time timeout 1500 python -m pytest tests -m gpu -x -q) > gpurun_out/r2_pytest_gpu_9.log 2>&1
tail -4 gpurun_out/r2_pytest_gpu_9.log
