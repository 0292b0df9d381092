#!/bin/bash
cd "$(dirname "$0")/.."
out=gpurun_out/r2_run3.log
: > $out
echo "== smoke" >> $out
timeout 300 python __graft_entry__.py smoke >> $out 2>&1; echo "smoke rc=$?" >> $out
echo "== default lib, shapes" >> $out
EOSVR_EXP=16 timeout 300 python tools/shape_perf.py cfg-3 cfg-2 cfg-4 >> $out 2>&1; echo "rc=$?" >> $out
export EOSVR_LIB_PATH=$PWD/embodied-one-shot-video-recognition_b200/libeosvr_exp.so
for ex in 48 54; do
  for ew in 16 8; do
    echo "== EXP=$ex EW=$ew (16 profile, +4 no MMA, +2 no TMA, +32 no rare path)" >> $out
    EOSVR_SELFCHECK=0 EOSVR_EW=$ew EOSVR_EXP=$ex timeout 200 python tools/shape_perf.py "cfg-3 E=1024" >> $out 2>&1; echo "rc=$?" >> $out
  done
done
unset EOSVR_LIB_PATH
(time timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/r2_pytest_gpu_2.log 2>&1
tail -4 gpurun_out/r2_pytest_gpu_2.log >> $out
