#!/bin/bash
# A/B of lanes-per-bubble configurations: profiles/gl_sweep.sh "1,1,4,4,8" "1,1,2,4,8" ...
for cfg in "$@"; do
  PF_GROUP_LANES=$cfg python bench.py --no-cpu-baseline --steps 5 --warmup 3 $BENCH_ARGS 2> gpurun_out/gl.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('lanes','$cfg','value %.3g'%d['value'],'align ms %.2f'%d['ms_align_pipeline'],'e2e ms %.2f'%d['e2e']['ms_per_step'])
"
done
