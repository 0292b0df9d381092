#!/bin/bash
# early accumulator release in the aligned epilogue: knob parity + A/B against the previous commit's build
cd "$(dirname "$0")/.."
P=embodied-one-shot-video-recognition_b200
(time timeout 600 python -m pytest tests/test_gpu_orders.py -x -q) > gpurun_out/r2_pytest_orders_16.log 2>&1
tail -5 gpurun_out/r2_pytest_orders_16.log
timeout 900 python tools/ab_perf.py $P/libeosvr_prev.so $P/libeosvr.so 2 > gpurun_out/r2_ab_early_release.log 2>&1
cat gpurun_out/r2_ab_early_release.log
