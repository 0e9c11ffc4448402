#!/bin/bash
# A/B of the sub-step kernel variants of round 2 (run on the GPU box through gpurun)
cd "$(dirname "$0")/.."
run() { # name lib lenbits lent particles
  FPB_ENGINE_LIB=$PWD/flexpart_b200/$2 FPB_LEN_BITS=$3 FPB_LEN_T=$4 python bench.py --steps 12 --warmup 3 --no-cpu --no-hbm-regime --no-c5 --particles $5 \
    > gpurun_out/ab_$1.json 2> gpurun_out/ab_$1.err
  python - "$1" <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.load(open(f"gpurun_out/ab_{n}.json"))
    r = d["roofline"]
    print(f"{n:28s} value {d['value']:.4e}  ms/step {d['ms_per_step']:.4f}  step-kernels {r['kernel_ms_per_launch']:.4f}  conc {r['conccalc_ms_per_launch']:.4f}  e2e {d['e2e']['value']:.3e}")
except Exception as e:
    print(n, "FAILED", e)
PY
}
run base_1m       libfpb_nopf.so 0 48 1000000
run lpt2_1m       libfpb_nopf.so 2 48 1000000
run pf_1m         libfpb.so      0 48 1000000
run pf_lpt2_1m    libfpb.so      2 48 1000000
run pf_lpt1_1m    libfpb.so      1 32 1000000
run pf_lpt2t64_1m libfpb.so      2 64 1000000
run pf_lpt2t32_1m libfpb.so      2 32 1000000
run pf_lpt3_1m    libfpb.so      3 64 1000000
run base_8m       libfpb_nopf.so 0 48 8000000
run pf_lpt2_8m    libfpb.so      2 48 8000000
run pf_8m         libfpb.so      0 48 8000000
