#!/bin/bash
cd "$(dirname "$0")/.."
out=gpurun_out/r2_run6.log
: > $out
echo "== smoke" >> $out
timeout 300 python __graft_entry__.py smoke >> $out 2>&1; echo "smoke rc=$?" >> $out
for st in 2 4 8 16; do
  echo "== production kernels, EOSVR_SEED_TILES=$st" >> $out
  EOSVR_SEED_TILES=$st timeout 300 python tools/shape_perf.py "cfg-3 E=1024" cfg-2 cfg-4 >> $out 2>&1; echo "rc=$?" >> $out
done
echo "== EXP=16 (diag kernels) seed 8" >> $out
EOSVR_SEED_TILES=8 EOSVR_EXP=16 timeout 300 python tools/shape_perf.py "cfg-3 E=1024" cfg-2 >> $out 2>&1; echo "rc=$?" >> $out
export EOSVR_LIB_PATH=$PWD/embodied-one-shot-video-recognition_b200/libeosvr_exp.so
for ex in 32 38; do
    echo "== EXP=$ex EW=16 no profile" >> $out
    EOSVR_SELFCHECK=0 EOSVR_EXP=$ex timeout 200 python tools/shape_perf.py "cfg-3 E=1024" >> $out 2>&1; echo "rc=$?" >> $out
done
unset EOSVR_LIB_PATH
(time timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/r2_pytest_gpu_4.log 2>&1
tail -4 gpurun_out/r2_pytest_gpu_4.log >> $out
