#!/bin/bash
# Run the bring-up cases one by one, each under its own timeout, logging to gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
for c in "$@"; do
  timeout 240 python tools/gpu_check.py --case $c > gpurun_out/check_$c.log 2>&1
  echo "case $c exit $?" | tee -a gpurun_out/summary.txt
  tail -n 25 gpurun_out/check_$c.log
done
