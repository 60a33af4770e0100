#!/bin/bash
# quick developer loop on the GPU box: debug counters, per-kernel times (both refine table placements), parity on 1 M states
mkdir -p gpurun_out
PBP_ENGINE_LIB=pyroboplan_b200/libpbp_engine_dbg.so python tools/gpu_dbg_stats.py 2>&1 | tee gpurun_out/quick_dbg.log | grep -v "^ *[0-9]* *0$"
echo "--- product build, per-kernel ms over 5 rounds"
python tools/gpu_dbg_stats.py noprint 2>&1 | tail -1
echo "--- PBP_SR_SMEM=0"
PBP_SR_SMEM=0 python tools/gpu_dbg_stats.py noprint 2>&1 | tail -1
if [ "$1" = "parity" ]; then timeout 600 python tools/gpu_surface_check.py 1000000 0 0.0 2>&1 | tail -7; fi
