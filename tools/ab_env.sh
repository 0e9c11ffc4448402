#!/bin/bash
# A/B of run-time knobs (run on the GPU box through gpurun): tools/ab_env.sh "VAR=a" "VAR=b" ...
cd "$(dirname "$0")/.."
for kv in "$@"; do
  for T in 12500000 100000000; do
    env $kv FPB_C5_TOTAL=$T python bench.py --workload c5 --steps 16 --warmup 3 2>/dev/null | python -c "
import json,sys; c=json.loads(sys.stdin.read().strip().splitlines()[-1])['c5_strong']
print('$kv c5 $T: %.4g  %.3f ms/step  kernels %.3f  frac %.3f' % (c['value'], c['ms_per_step'], c['kernel_ms_per_launch'], c['roofline']['frac']))"
  done
  env $kv python bench.py --steps 12 --warmup 3 --no-cpu --no-c5 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$kv C2: %.4g  %.3f ms/step  kernels %.3f;  hbm_regime %.4g frac %.3f' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms_per_launch'], d['hbm_regime']['value'], d['hbm_regime']['roofline']['frac']))"
done
