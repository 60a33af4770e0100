#!/bin/bash
# A/B timing of tuning builds of libpbp_engine.so (same ABI): run through gpurun
for lib in "" build_variants/libpbp_*.so; do
  echo "=== ${lib:-default}"
  if [ -n "$lib" ]; then export PBP_ENGINE_LIB=$PWD/$lib; else unset PBP_ENGINE_LIB; fi
  python tools/gpu_size_sweep.py 2>&1 | grep "N= 1048576\|N=  262144"
done
