#!/bin/bash
# configs 4 and 5 and the headline bench on N GPUs of one box (N = $1): gpurun --gpus N --timeout 1500 -- 'bash tools/gpu_multi.sh N'
N=${1:-2}
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
if [ "$N" = "1" ]; then RUN="python"; fi
timeout 600 $RUN tools/gpu_config4.py 2> gpurun_out/r02_config4_n$N.err | tail -1 > gpurun_out/r02_config4_n$N.json; cat gpurun_out/r02_config4_n$N.json
timeout 600 $RUN tools/gpu_ik_sweep.py 2> gpurun_out/r02_ik_n$N.err | tail -1 > gpurun_out/r02_ik_n$N.json; cat gpurun_out/r02_ik_n$N.json
timeout 600 $RUN bench.py --gpus $N --steps 20 --warmup 3 2> gpurun_out/r02_bench_n$N.err | tail -1 > gpurun_out/r02_bench_n$N.json
python - <<PY
import json
d=json.loads(open('gpurun_out/r02_bench_n$N.json').read())
print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, d['e2e']['value'], d['e2e']['pageable_f64']['value'])
PY
