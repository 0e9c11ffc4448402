#!/bin/bash
# A/B of engine build variants (run on the GPU box through gpurun): C5 at 12.5 M and 100 M particles, C2
cd "$(dirname "$0")/.."
run() {
  name=$1; lib=$2
  for T in 12500000 100000000; do
    FPB_C5_TOTAL=$T FPB_ENGINE_LIB=$PWD/flexpart_b200/$lib python bench.py --workload c5 --steps 16 --warmup 3 2>/dev/null | python -c "
import json,sys; c=json.loads(sys.stdin.read().strip().splitlines()[-1])['c5_strong']
print('$name c5 $T: %.4g  %.3f ms/step  kernels %.3f  frac %.3f' % (c['value'], c['ms_per_step'], c['kernel_ms_per_launch'], c['roofline']['frac']))"
  done
  FPB_ENGINE_LIB=$PWD/flexpart_b200/$lib python bench.py --steps 12 --warmup 3 --no-cpu --no-c5 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$name C2: %.4g  %.3f ms/step  kernels %.3f;  hbm_regime %.4g frac %.3f' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms_per_launch'], d['hbm_regime']['value'], d['hbm_regime']['roofline']['frac']))"
}
for v in "$@"; do run $v libfpb_$v.so; done
run base libfpb.so
