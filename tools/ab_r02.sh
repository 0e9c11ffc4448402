#!/bin/bash
# A/B of engine build variants on the C5 workload (run on the GPU box through gpurun)
cd "$(dirname "$0")/.."
run() {
  name=$1; lib=$2; shift 2
  FPB_ENGINE_LIB=$PWD/flexpart_b200/$lib python bench.py --workload c5 --steps 6 --warmup 3 "$@" > gpurun_out/ab_$name.json 2> gpurun_out/ab_$name.err
  python - "$name" <<'PY'
import json, sys
n = sys.argv[1]
try:
    c = json.load(open(f"gpurun_out/ab_{n}.json"))["c5_strong"]
    print(f"{n:14s} c5 {c['value']:.4e} ms/step {c['ms_per_step']:.3f} kernels {c['kernel_ms_per_launch']:.3f} conc {c['conccalc_ms_per_launch']:.3f} frac {c['roofline']['frac']:.3f}")
except Exception as e:
    print(n, "FAILED", e)
PY
}
run fb8  libfpb.so
run fb4  libfpb_fb4.so
run fb6  libfpb_fb6.so
run fb10 libfpb_fb10.so
run fb12 libfpb_fb12.so
