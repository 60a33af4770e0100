#!/bin/bash
# gpurun --timeout 1800 -- 'bash tools/gpu_round2b.sh'
mkdir -p gpurun_out
timeout 600 python tools/gpu_surface_check.py 1000000 0 0.0 > gpurun_out/r2b_surface_check.log 2>&1; tail -7 gpurun_out/r2b_surface_check.log
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r2b_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2b_bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','parity')}); print(d['e2e']['value'], d['e2e']['pageable_f64']['value']); print(d['roofline']['kernel_ms'], d['roofline']['frac'])
PY
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2b_pytest_gpu.log 2>&1; tail -25 gpurun_out/r2b_pytest_gpu.log
