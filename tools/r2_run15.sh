#!/bin/bash
# cycle accounting of the episode-aligned screening kernel on cfg-3 (experiments build: EOSVR_EXP 16 = accounting,
# +32 no candidate handling, +4 no MMA, +2 no TMA, +1 no epilogue)
cd "$(dirname "$0")/.."
out=gpurun_out/r2_run15.log
: > $out
export EOSVR_LIB_PATH=$PWD/embodied-one-shot-video-recognition_b200/libeosvr_exp.so
for ex in 48 54 17; do
    echo "== EXP=$ex aligned" >> $out
    EOSVR_SELFCHECK=0 EOSVR_EXP=$ex timeout 200 python tools/shape_perf.py "cfg-3 E=1024" >> $out 2>&1; echo "rc=$?" >> $out
done
echo "== EXP=16 generic (EOSVR_ALIGNED=0)" >> $out
EOSVR_ALIGNED=0 EOSVR_EXP=16 timeout 200 python tools/shape_perf.py "cfg-3 E=1024" >> $out 2>&1; echo "rc=$?" >> $out
cat $out
