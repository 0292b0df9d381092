#!/bin/bash
cd "$(dirname "$0")/.."
out=gpurun_out/r2_run7.log
: > $out
timeout 300 python __graft_entry__.py smoke >> $out 2>&1; echo "smoke rc=$?" >> $out
for od in 1 0 2; do
  echo "== EOSVR_ORDER=$od" >> $out
  EOSVR_ORDER=$od timeout 300 python tools/shape_perf.py "cfg-3 E=1024" cfg-2 cfg-4 cfg-5 >> $out 2>&1; echo "rc=$?" >> $out
done
for tpu in 4 8 32; do
  echo "== EOSVR_ORDER=0 EOSVR_TPU=$tpu" >> $out
  EOSVR_ORDER=0 EOSVR_TPU=$tpu timeout 300 python tools/shape_perf.py "cfg-3 E=1024" cfg-4 >> $out 2>&1; echo "rc=$?" >> $out
done
(time timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/r2_pytest_gpu_5.log 2>&1
tail -4 gpurun_out/r2_pytest_gpu_5.log >> $out
