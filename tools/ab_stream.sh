#!/bin/bash
# A/B of the streamed fpb_step_host (one persistent sub-step launch for all chunks) against the
# chunk-by-chunk pipeline; run on the GPU box:  bash tools/ab_stream.sh > gpurun_out/ab_stream.txt
cd "$(dirname "$0")/.."
run() {
  env FPB_HOST_STREAM=1 "$@" timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu --no-c5 --no-hbm-regime \
    2> /tmp/ab_stream.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$*  e2e %.4g particle-steps/s  (%.3f ms/step)  resident %.3f ms/step' % (d['e2e']['value'], 1e3*d['config']['particles_per_gpu']/d['e2e']['value'], d['ms_per_step']))
" || tail -3 /tmp/ab_stream.err
}
run FPB_HOST_STREAM=0
run FPB_HOST_GRID_FRAC=0.6
for fr in ${FRACS:-0.5 0.7}; do
  run FPB_HOST_GRID_FRAC=$fr
done
for plan in ${PLANS:-0.1,0.3,0.3,0.2,0.07,0.03 0.05,0.1,0.15,0.2,0.2,0.15,0.08,0.04,0.03 0.08,0.25,0.25,0.2,0.12,0.06,0.04}; do
  run FPB_HOST_PLAN=$plan
done
for ch in ${CHUNKS:-4 8}; do
  run FPB_HOST_CHUNKS=$ch
done
