#!/bin/bash
# chunk plans / copy-out order of the fpb_step_host pipeline on C2; run on the GPU box
cd "$(dirname "$0")/.."
run() {
  env "$@" timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu --no-c5 --no-hbm-regime \
    2> /tmp/ab_plan.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$*  e2e %.4g particle-steps/s  (%.3f ms/step)  resident %.3f ms/step' % (d['e2e']['value'], 1e3*d['config']['particles_per_gpu']/d['e2e']['value'], d['ms_per_step']))
" || tail -3 /tmp/ab_plan.err
}
run FPB_HOST_X=default
run FPB_HOST_PLAN=equal FPB_HOST_DEFER_D2H=0
run FPB_HOST_PLAN=equal
run FPB_HOST_DEFER_D2H=0
for plan in ${PLANS:-0.3,0.3,0.25,0.15 0.28,0.28,0.24,0.2 0.3,0.3,0.3,0.1}; do
  run FPB_HOST_PLAN=$plan
done
run FPB_HOST_X=default
