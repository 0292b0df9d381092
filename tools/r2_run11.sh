#!/bin/bash
cd "$(dirname "$0")/.."
out=gpurun_out/r2_run11.log
: > $out
export EOSVR_LIB_PATH=$PWD/embodied-one-shot-video-recognition_b200/libeosvr_exp.so
for ex in 17 38 32 54; do
    echo "== EXP=$ex EW=16" >> $out
    EOSVR_SELFCHECK=0 EOSVR_EXP=$ex timeout 200 python tools/shape_perf.py "cfg-3 E=1024" cfg-5 >> $out 2>&1; echo "rc=$?" >> $out
done
