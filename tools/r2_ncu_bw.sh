#!/bin/bash
# ncu --set full of the bandwidth kernels inside the bench's cfg-3 step (one launch each), after the same command ran plain
cd "$(dirname "$0")/.."
python bench.py --profile --steps 6 --warmup 3 > gpurun_out/r2_ncu_bw_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"k_rerank_rows|k_episode_partial|k_probe_prep|k_finish" -s 28 -c 4 -f -o gpurun_out/r2_prof_bw_cfg3 \
    python bench.py --profile --steps 6 --warmup 3 > gpurun_out/r2_ncu_bw_run.log 2>&1
echo "rc=$?" >> gpurun_out/r2_ncu_bw_run.log
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 40 --csv --log-file gpurun_out/r2_launches_bench_profile.csv \
    python bench.py --profile --steps 6 --warmup 3 > gpurun_out/r2_ncu_list.log 2>&1
