#!/bin/bash
# per-chunk timeline of fpb_step_host on the bench's C2 workload (FPB_HOST_TIMING=1), last timed call only
cd "$(dirname "$0")/.."
tl() {
  echo "== $*"
  env FPB_HOST_TIMING=1 "$@" timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu --no-c5 --no-hbm-regime 2>&1 >/dev/null \
    | grep "fpb_step_host" | tail -${TL_LINES:-12}
}
tl FPB_HOST_X=0
tl FPB_HOST_PLAN=0.3,0.3,0.25,0.15
tl FPB_HOST_CHUNKS=8
