#!/bin/bash
# ncu --set full of the screening kernel (episode-aligned epilogue) inside the bench's cfg-3 step
cd "$(dirname "$0")/.."
python bench.py --profile --steps 6 --warmup 3 > gpurun_out/r2_ncu_bench2_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_match_screen -s 14 -c 2 -f -o gpurun_out/r2_prof_bench2_cfg3 \
    python bench.py --profile --steps 6 --warmup 3 > gpurun_out/r2_ncu_bench2_run.log 2>&1
echo "rc=$?" >> gpurun_out/r2_ncu_bench2_run.log
tail -3 gpurun_out/r2_ncu_bench2_run.log
