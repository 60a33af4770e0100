#!/bin/bash
# Regenerates every measured artefact of a round in ONE gpurun call (copy the results into profiles/):
#   gpurun --timeout 1800 -- 'bash tools/gpu_round1.sh'
# smoke, GPU test suite, bench (ours + reference arm), per-config timings, single-query latency, surface-semantics parity,
# ncu launch list + full capture of one round's three kernels.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; tail -1 gpurun_out/pytest_gpu.log
python bench.py > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; echo "rc=$?" >> gpurun_out/bench_ours.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
python tools/gpu_configs_bench.py > gpurun_out/configs_bench.json 2>&1
python tools/gpu_latency.py > gpurun_out/latency.log 2>&1
python tools/gpu_surface_check.py 1000000 0 0.0 > gpurun_out/surface_check.log 2>&1   # engine vs the oracle's triangle tests, both mesh semantics
python tools/gpu_surface_sweep.py 10 4 1000000 > gpurun_out/surface_sweep.log 2>&1   # verdict parity on more seeds
bash tools/gpu_ncu.sh
ls -la gpurun_out | tail -20
