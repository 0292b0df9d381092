#!/bin/bash
# read-ahead exact chains + k_episode_dist: parity subset, per-kernel times of the cfg-3 / cfg-2 step
cd "$(dirname "$0")/.."
T=${TAG:-c3}
(time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_dropin.py -x -q) > gpurun_out/r2${T}_pytest.log 2>&1; tail -6 gpurun_out/r2${T}_pytest.log
EOSVR_EXP=64 timeout 600 python tools/step_profile.py > gpurun_out/r2${T}_step_profile_cfg3.log 2>&1; grep -v Warn gpurun_out/r2${T}_step_profile_cfg3.log
EOSVR_EXP=64 timeout 600 python tools/step_profile.py cfg2 > gpurun_out/r2${T}_step_profile_cfg2.log 2>&1; grep -v Warn gpurun_out/r2${T}_step_profile_cfg2.log
