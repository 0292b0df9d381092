#!/bin/bash
# Screening-kernel experiments on the bench workload (results are wrong for EXP bits 1/2/4).
for np in 1 2; do for ex in 16 17; do
  echo "NP=$np EXP=$ex"; EOSVR_NP=$np EOSVR_EXP=$ex timeout 200 python tools/gpu_check.py --case perf 2>&1 | grep -E "match P|cycles|Error|error" | cut -c1-250
done; done
echo "NP=1 clustered"; EOSVR_NP=1 EOSVR_EXP=16 timeout 200 python tools/gpu_check.py --case perf_clustered 2>&1 | grep -E "match P|cycles|Error|error" | cut -c1-250
echo "NP=2 clustered"; EOSVR_NP=2 EOSVR_EXP=16 timeout 200 python tools/gpu_check.py --case perf_clustered 2>&1 | grep -E "match P|cycles|Error|error" | cut -c1-250
