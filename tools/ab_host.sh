#!/bin/bash
# A/B of the fpb_step_host pipeline knobs (chunks x persistent-grid fraction); run on the GPU box:
#   bash tools/ab_host.sh > gpurun_out/ab_host.txt
cd "$(dirname "$0")/.."
for ch in 3 4 6 8 12; do
  for fr in 1 0.6 0.4 0.2; do
    FPB_HOST_CHUNKS=$ch FPB_HOST_GRID_FRAC=$fr python bench.py --steps 8 --warmup 3 --no-cpu --no-c5 --no-hbm-regime \
      2> /tmp/ab_host.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('chunks $ch frac $fr  e2e %.4g particle-steps/s  (%.3f ms/step)  resident %.3f ms/step' % (d['e2e']['value'], 1e3*d['config']['particles_per_gpu']/d['e2e']['value'], d['ms_per_step']))
"
  done
done
