#!/bin/bash
# ncu capture of the hot kernels (one gpurun call; plain run first as the profiling guide requires)
mkdir -p gpurun_out
export PBP_CPU_SAMPLE=20000
python bench.py --steps 2 --warmup 3 > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"k_bp_spec|k_broadphase|k_item_setup|k_narrowphase|k_surface_refine|k_surface_settle" -s 15 -c 5 -o gpurun_out/prof_round python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_round.log 2>&1
python bench.py --steps 2 --warmup 3 > gpurun_out/plain2.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_list.log 2>&1
tail -3 gpurun_out/ncu_round.log
