#!/bin/bash
# A/B of engine build variants on C2 (1 M and 8 M particles) and C3 (run on the GPU box through gpurun)
cd "$(dirname "$0")/.."
run() {
  name=$1; lib=$2
  for P in 1000000 8000000; do
    FPB_ENGINE_LIB=$PWD/flexpart_b200/$lib python bench.py --steps 12 --warmup 3 --no-cpu --no-c5 --no-hbm-regime --particles $P 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$name C2 $P: %.4g  %.3f ms/step  kernels %.3f  e2e %.4g' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms_per_launch'], d['e2e']['value']))"
  done
}
for v in "$@"; do run $v libfpb_$v.so; done
run base libfpb.so
