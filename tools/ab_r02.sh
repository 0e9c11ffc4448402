#!/bin/bash
# A/B of engine build variants (run on the GPU box through gpurun): name lib particles [extra bench args]
cd "$(dirname "$0")/.."
run() {
  name=$1; lib=$2; n=$3; shift 3
  FPB_ENGINE_LIB=$PWD/flexpart_b200/$lib python bench.py --steps 12 --warmup 3 --no-cpu --no-hbm-regime --particles $n "$@" \
    > gpurun_out/ab_$name.json 2> gpurun_out/ab_$name.err
  python - "$name" <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.load(open(f"gpurun_out/ab_{n}.json"))
    r = d["roofline"]
    s = f"{n:22s} value {d['value']:.4e} ms/step {d['ms_per_step']:.4f} step-kernels {r['kernel_ms_per_launch']:.4f} conc {r['conccalc_ms_per_launch']:.4f} e2e {d['e2e']['value']:.3e}"
    c = d.get("c5_strong")
    if c:
        s += f" | c5 {c['value']:.4e} ms {c['ms_per_step']:.3f} kern {c['kernel_ms_per_launch']:.3f} conc {c['conccalc_ms_per_launch']:.3f}"
    print(s)
except Exception as e:
    print(n, "FAILED", e)
PY
}
run base_1m   libfpb.so      1000000
run noagg_1m  libfpb_noagg.so 1000000
run sr4_1m    libfpb_sr4.so  1000000 --no-c5
run sr8_1m    libfpb_sr8.so  1000000 --no-c5
run base_8m   libfpb.so      8000000 --no-c5
run sr8_8m    libfpb_sr8.so  8000000 --no-c5
