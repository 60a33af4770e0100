#!/bin/bash
# One gpurun call: smoke, bench (ours + reference arm), ncu launch list, ncu full capture of the two hot kernels.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; echo "rc=$?" >> gpurun_out/bench_ours.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
export PBP_CPU_SAMPLE=20000
python bench.py --steps 2 --warmup 3 > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_list.log 2>&1
python bench.py --steps 2 --warmup 3 > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_narrowphase -s 3 -c 2 -o gpurun_out/prof_narrowphase python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_np.log 2>&1
python bench.py --steps 2 --warmup 3 > gpurun_out/plain3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_broadphase -s 3 -c 2 -o gpurun_out/prof_broadphase python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_bp.log 2>&1
ls -la gpurun_out
