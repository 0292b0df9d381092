#!/bin/bash
# round-2 evidence on one B200 with the final build: GPU test-suite, smoke, bench (both arms), ncu launch list and
# ncu --set full captures (each only after the same command ran plain)
cd "$(dirname "$0")/.."
T=${TAG:-final}
(time timeout 1200 python -m pytest tests -m gpu -x -q) > gpurun_out/r2_${T}_pytest_gpu.log 2>&1; tail -3 gpurun_out/r2_${T}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_${T}_smoke.log 2>&1; tail -1 gpurun_out/r2_${T}_smoke.log
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_${T}_bench_reference.json 2> gpurun_out/r2_${T}_bench_reference.err; echo "ref rc=$?"
timeout 1200 python bench.py > gpurun_out/r2_${T}_bench_n1.json 2> gpurun_out/r2_${T}_bench_n1.err; echo "bench rc=$?"
cut -c1-600 gpurun_out/r2_${T}_bench_n1.json
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit,clocks_throttle_reasons.active --format=csv > gpurun_out/r2_${T}_smi.txt 2>&1
python bench.py --profile --steps 4 --warmup 3 > gpurun_out/r2_${T}_profile_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_${T}_launches.csv \
    python bench.py --profile --steps 4 --warmup 3 > gpurun_out/r2_${T}_ncu_list.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"k_match_screen|k_rerank|k_episode_partial|k_probe_prep|k_finish" -s 30 -c 6 -f \
    -o gpurun_out/r2_${T}_prof_step python bench.py --profile --steps 4 --warmup 3 > gpurun_out/r2_${T}_ncu_full.log 2>&1
echo "ncu full rc=$?"
