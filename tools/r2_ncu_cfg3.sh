#!/bin/bash
# ncu --set full of the screening kernel on cfg-3 (E=1024, D=512, G=100k); the same command first runs without ncu
cd "$(dirname "$0")/.."
python tools/shape_perf.py "cfg-3 E=1024" > gpurun_out/r2_ncu_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_match_screen -s 8 -c 2 -f -o gpurun_out/r2_prof_cfg3 \
    python tools/shape_perf.py "cfg-3 E=1024" > gpurun_out/r2_ncu_run.log 2>&1
echo "rc=$?" >> gpurun_out/r2_ncu_run.log
