#!/bin/bash
# same-box A/B of engine builds on C2 / the HBM regime: libfpb.so against variants built by tools/build_variant.sh
# (flexpart_b200/libfpb_TAG.so); usage: bash tools/ab_pblloop.sh [TAG ...].  profiles/ab_r02_pblloop.txt was made with
# it (base = the sub-step loop with the streamed-host-step tests, orig = the loop as it is now)
cd "$(dirname "$0")/.."
run() {
  name=$1; lib=$2
  for rep in 1 2; do
  FPB_ENGINE_LIB=$PWD/flexpart_b200/$lib python bench.py --steps 24 --warmup 4 --no-cpu --no-c5 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$name C2: %.4g  %.3f ms/step  kernels %.3f;  hbm_regime %.4g frac %.3f kernels %.3f' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms_per_launch'], d['hbm_regime']['value'], d['hbm_regime']['roofline']['frac'], d['hbm_regime']['kernel_ms_per_launch']))"
  done
}
run base libfpb.so
for v in "$@"; do run $v libfpb_$v.so; done
run base libfpb.so
