#!/bin/bash
cd "$(dirname "$0")/.."
out=gpurun_out/r2_run4.log
: > $out
export EOSVR_LIB_PATH=$PWD/embodied-one-shot-video-recognition_b200/libeosvr_exp.so
# 54 = 16|2|4|32: profile, no TMA, no MMA, no rare path.  +128 no refill, +256 no quadrant barrier, +512 no sqrt
for ex in 54 182 310 566 950; do
  for ew in 16; do
    echo "== EXP=$ex EW=$ew" >> $out
    EOSVR_SELFCHECK=0 EOSVR_EW=$ew EOSVR_EXP=$ex timeout 200 python tools/shape_perf.py "cfg-3 E=1024" >> $out 2>&1; echo "rc=$?" >> $out
  done
done
