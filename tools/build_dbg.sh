#!/bin/bash
# developer build with the kernels' debug counters (tools/gpu_dbg_stats.py); extra -D flags as arguments
cd "$(dirname "$0")/.."
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -Xcompiler -fPIC -shared -ldl -DPBP_DEBUG_STATS "$@" \
    -o pyroboplan_b200/libpbp_engine_dbg.so pyroboplan_b200/csrc/pbp_engine.cu
