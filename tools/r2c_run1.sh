#!/bin/bash
# scipy-order exact evaluation (k_exact_jobs, lane-per-pair exhaustive kernel, thread-per-pair prototype distances):
# the new order-sensitive / accumulation-term tests first, then the whole GPU suite, smoke, and a primary-only bench
cd "$(dirname "$0")/.."
T=${TAG:-c1}
(time timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -s -k "order_sensitive or accumulation or match_vs_oracle or golden_classifier or candidate_overflow or one_call") > gpurun_out/r2${T}_pytest_new.log 2>&1; tail -8 gpurun_out/r2${T}_pytest_new.log
(time timeout 1500 python -m pytest tests -m gpu -q) > gpurun_out/r2${T}_pytest_gpu.log 2>&1; tail -8 gpurun_out/r2${T}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2${T}_smoke.log 2>&1; tail -1 gpurun_out/r2${T}_smoke.log
timeout 900 python bench.py --primary-only --no-cpu > gpurun_out/r2${T}_bench_primary.json 2> gpurun_out/r2${T}_bench_primary.err; echo "bench rc=$?"
cut -c1-300 gpurun_out/r2${T}_bench_primary.json
