#!/bin/bash
# chunk plans of the (chunk-by-chunk) fpb_step_host pipeline on C2; run on the GPU box
cd "$(dirname "$0")/.."
run() {
  env "$@" timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu --no-c5 --no-hbm-regime \
    2> /tmp/ab_plan.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$*  e2e %.4g particle-steps/s  (%.3f ms/step)  resident %.3f ms/step' % (d['e2e']['value'], 1e3*d['config']['particles_per_gpu']/d['e2e']['value'], d['ms_per_step']))
" || tail -3 /tmp/ab_plan.err
}
run FPB_HOST_X=0
for plan in ${PLANS:-0.3,0.3,0.25,0.15 0.3,0.3,0.25,0.1,0.05 0.35,0.35,0.2,0.07,0.03 0.4,0.3,0.2,0.1 0.25,0.25,0.25,0.15,0.07,0.03 0.5,0.3,0.15,0.05}; do
  run FPB_HOST_PLAN=$plan
done
for fr in 0.6 1.0; do run FPB_HOST_PLAN=0.3,0.3,0.25,0.1,0.05 FPB_HOST_GRID_FRAC=$fr; done
run FPB_HOST_DEFER_D2H=1
for plan in 0.3,0.3,0.25,0.15 0.3,0.3,0.25,0.1,0.05 0.4,0.3,0.2,0.1; do
  run FPB_HOST_DEFER_D2H=1 FPB_HOST_PLAN=$plan
done
