#!/bin/bash
# e2e slice-size sweep for the host-buffer path (run through gpurun)
for s in 131072 262144 393216; do
  PBP_HOST_SLICE=$s python bench.py --steps 100 --warmup 10 2> gpurun_out/sweep_$s.err | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print($s, 'value %.3e e2e %.3e ms/step %.3f' % (d['value'], d['e2e']['value'], d['ms_per_step']), d['clocks'], d['parity'])"
done
