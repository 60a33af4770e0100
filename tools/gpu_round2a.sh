#!/bin/bash
# Round 2 first check of the exact surface semantics on the GPU (one gpurun call):
#   gpurun --timeout 1500 -- 'bash tools/gpu_round2a.sh'
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2a_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2a_smoke.log
tail -3 gpurun_out/r2a_smoke.log
timeout 600 python tools/gpu_surface_check.py 1000000 0 0.0 > gpurun_out/r2a_surface_check.log 2>&1; tail -12 gpurun_out/r2a_surface_check.log
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench rc=$?"; cat gpurun_out/r2a_bench.json | head -c 3000
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2a_pytest_gpu.log 2>&1; tail -15 gpurun_out/r2a_pytest_gpu.log
