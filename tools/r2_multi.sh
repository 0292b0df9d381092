#!/bin/bash
# multi-GPU parity + strong-scaling bench on N GPUs of one box:  bash tools/r2_multi.sh N
cd "$(dirname "$0")/.."
N=${1:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    tools/multi_gpu_check.py > gpurun_out/r2_multi_gpu_check_n$N.txt 2> gpurun_out/r2_multi_gpu_check_n$N.err
echo "multi_gpu_check rc=$?"; cat gpurun_out/r2_multi_gpu_check_n$N.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err
echo "bench rc=$?"; tail -3 gpurun_out/r2_bench_n$N.err
timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -x -q > gpurun_out/r2_pytest_multi_n$N.log 2>&1; tail -2 gpurun_out/r2_pytest_multi_n$N.log
