#!/bin/bash
# Race hunting: repeated matches for both cluster shapes and several sizes; stops at the first failure/timeout.
for np in 1 2; do for cfg in "300 64 2048" "150 256 2048" "300 40 64" "300 40 192" "300 40 512"; do
  echo "NP=$np cfg=$cfg"
  EOSVR_NP=$np timeout 40 python tools/stress_match.py $cfg 2>&1 | tail -1
  rc=${PIPESTATUS[0]}
  if [ $rc -ne 0 ]; then echo "STOP rc=$rc"; exit 1; fi
done; done
echo ALL-OK
