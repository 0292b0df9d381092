#!/bin/bash
# ncu launch list + one --set full capture of every kernel of the bench's cfg-3 step (after the same command ran plain)
cd "$(dirname "$0")/.."
T=${TAG:-final}
python bench.py --profile --steps 4 --warmup 3 > gpurun_out/r2_${T}_profile_plain.log 2>&1 || { tail -5 gpurun_out/r2_${T}_profile_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_${T}_launches.csv \
    python bench.py --profile --steps 4 --warmup 3 > gpurun_out/r2_${T}_ncu_list.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"k_match_screen|k_rerank|k_episode_partial|k_probe_prep|k_finish" -s 30 -c 6 -f \
    -o gpurun_out/r2_${T}_prof_step python bench.py --profile --steps 4 --warmup 3 > gpurun_out/r2_${T}_ncu_full.log 2>&1
echo "ncu full rc=$?"
