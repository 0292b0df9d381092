#!/bin/bash
# accumulation-term test on full-mantissa data; per-kernel times of the cfg-3 / cfg-2 step (k_exact_jobs); default bench duration
cd "$(dirname "$0")/.."
T=${TAG:-c2}
(time timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -s -k "accumulation") > gpurun_out/r2${T}_pytest_acc.log 2>&1; tail -8 gpurun_out/r2${T}_pytest_acc.log
EOSVR_EXP=64 timeout 600 python tools/step_profile.py > gpurun_out/r2${T}_step_profile_cfg3.log 2>&1; cat gpurun_out/r2${T}_step_profile_cfg3.log
EOSVR_EXP=64 timeout 600 python tools/step_profile.py cfg2 > gpurun_out/r2${T}_step_profile_cfg2.log 2>&1; cat gpurun_out/r2${T}_step_profile_cfg2.log
(time timeout 1500 python bench.py > gpurun_out/r2${T}_bench_n1.json 2> gpurun_out/r2${T}_bench_n1.err) 2>&1 | tail -4; echo "bench rc=$?"
cut -c1-200 gpurun_out/r2${T}_bench_n1.json
