#!/bin/bash
cd "$(dirname "$0")/.."
P=embodied-one-shot-video-recognition_b200
timeout 1500 python tools/ab_perf.py $P/libeosvr_prev.so $P/libeosvr.so 2 > gpurun_out/r2_ab_normslab.log 2>&1
cat gpurun_out/r2_ab_normslab.log
