#!/bin/bash
# last build of round 2: whole GPU suite, smoke, primary bench line
cd "$(dirname "$0")/.."
T=${TAG:-c11}
(time timeout 1500 python -m pytest tests -m gpu -q) > gpurun_out/r2${T}_pytest_gpu.log 2>&1; tail -4 gpurun_out/r2${T}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2${T}_smoke.log 2>&1; tail -1 gpurun_out/r2${T}_smoke.log
timeout 900 python bench.py --primary-only --no-cpu > gpurun_out/r2${T}_bench_primary.json 2> gpurun_out/r2${T}_bench_primary.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2c11_bench_primary.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['parity_ok'], d['matcher_stats'], {k:round(v['ms'],4) for k,v in d['roofline']['kernels'].items()})
PY
