#!/bin/bash
# Screening-kernel experiments on the bench workload (results are wrong for EXP bits 1/2/4/32).
for np in ${NPS:-1 2}; do for ex in ${EXPS:-16}; do for od in ${ORDERS:-0}; do
  echo "NP=$np EXP=$ex ORDER=$od TPU=${EOSVR_TPU:-auto}"; EOSVR_ORDER=$od EOSVR_NP=$np EOSVR_EXP=$ex timeout 100 python tools/gpu_check.py --case perf 2>&1 | grep -E "match P|cycles|Error|error" | cut -c1-250
  echo "NP=$np EXP=$ex ORDER=$od clustered"; EOSVR_ORDER=$od EOSVR_NP=$np EOSVR_EXP=$ex timeout 100 python tools/gpu_check.py --case perf_clustered 2>&1 | grep -E "match P|cycles|Error|error" | cut -c1-250
done; done; done
