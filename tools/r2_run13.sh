#!/bin/bash
# aligned epilogue: knob parity first, then A/B against the generic epilogue of the same build, then the whole GPU suite
cd "$(dirname "$0")/.."
P=embodied-one-shot-video-recognition_b200
(time timeout 500 python -m pytest tests/test_gpu_orders.py -x -q) > gpurun_out/r2_pytest_orders_13.log 2>&1
tail -15 gpurun_out/r2_pytest_orders_13.log
timeout 600 python tools/ab_perf.py $P/libeosvr.so:EOSVR_ALIGNED=0 $P/libeosvr.so:EOSVR_ALIGNED=1 2 > gpurun_out/r2_ab_aligned.log 2>&1
cat gpurun_out/r2_ab_aligned.log
(time timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/r2_pytest_gpu_13.log 2>&1
tail -6 gpurun_out/r2_pytest_gpu_13.log
