#!/bin/bash
# three accumulator stages (160-column tiles) + one commit per stage pair: knob parity, then A/B against the previous commit
cd "$(dirname "$0")/.."
P=embodied-one-shot-video-recognition_b200
(time timeout 600 python -m pytest tests/test_gpu_orders.py -x -q) > gpurun_out/r2_pytest_orders_14.log 2>&1
tail -15 gpurun_out/r2_pytest_orders_14.log
timeout 900 python tools/ab_perf.py $P/libeosvr_prev.so $P/libeosvr.so:EOSVR_BN3=0 $P/libeosvr.so 2 > gpurun_out/r2_ab_bn3.log 2>&1
cat gpurun_out/r2_ab_bn3.log
(time timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/r2_pytest_gpu_14.log 2>&1
tail -6 gpurun_out/r2_pytest_gpu_14.log
