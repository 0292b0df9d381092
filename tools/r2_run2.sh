#!/bin/bash
# where does the in-situ epilogue time go?  (-DEOSVR_EXPERIMENTS build: EXP bits 1/2/4/32 give wrong results, timing only)
cd "$(dirname "$0")/.."
out=gpurun_out/r2_run2.log
export EOSVR_LIB_PATH=$PWD/embodied-one-shot-video-recognition_b200/libeosvr_exp.so
: > $out
for ex in 16 20 22 48 52 54 17; do
  for ew in 16 8; do
    echo "== EXP=$ex EW=$ew (16 profile, +4 no MMA, +2 no TMA, +32 no rare path, +1 no epilogue)" >> $out
    EOSVR_SELFCHECK=0 EOSVR_EW=$ew EOSVR_EXP=$ex timeout 200 python tools/shape_perf.py "cfg-3 E=1024" >> $out 2>&1; echo "rc=$?" >> $out
  done
done
