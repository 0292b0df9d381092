#!/bin/bash
# round-2 bring-up of the reworked screening kernel: parity first, then shapes
cd "$(dirname "$0")/.."
out=gpurun_out
echo "== smoke" > $out/r2_run1.log
timeout 300 python __graft_entry__.py smoke >> $out/r2_run1.log 2>&1; echo "smoke rc=$?" >> $out/r2_run1.log
echo "== shapes (default)" >> $out/r2_run1.log
timeout 300 python tools/shape_perf.py cfg-3 cfg-2 >> $out/r2_run1.log 2>&1; echo "rc=$?" >> $out/r2_run1.log
echo "== shapes EXP=16" >> $out/r2_run1.log
EOSVR_EXP=16 timeout 300 python tools/shape_perf.py "cfg-3 E=1024" "cfg-2" >> $out/r2_run1.log 2>&1; echo "rc=$?" >> $out/r2_run1.log
echo "== cfg-3 with EW=8" >> $out/r2_run1.log
EOSVR_EW=8 EOSVR_EXP=16 timeout 300 python tools/shape_perf.py "cfg-3 E=1024" >> $out/r2_run1.log 2>&1; echo "rc=$?" >> $out/r2_run1.log
echo "== cfg-2 with EW=16" >> $out/r2_run1.log
EOSVR_EW=16 EOSVR_EXP=16 timeout 300 python tools/shape_perf.py "cfg-2" >> $out/r2_run1.log 2>&1; echo "rc=$?" >> $out/r2_run1.log
echo "== cfg-4 shard" >> $out/r2_run1.log
timeout 300 python tools/shape_perf.py "cfg-4" "cfg-5" >> $out/r2_run1.log 2>&1; echo "rc=$?" >> $out/r2_run1.log
(time timeout 900 python -m pytest tests -m gpu -x -q) > $out/r2_pytest_gpu_1.log 2>&1
tail -5 $out/r2_pytest_gpu_1.log
