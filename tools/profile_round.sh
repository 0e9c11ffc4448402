#!/bin/bash
# tools/profile_round.sh ROUND -- run on the GPU box (gpurun -- 'bash tools/profile_round.sh r01'):
# the un-profiled bench lines first, then (same command, after it exited 0) the ncu launch list and
# one --set full capture of the three hot-path kernels.  Everything lands in gpurun_out/;
# tools/summarize_profiles.py turns it into profiles/.
set -u
R=${1:-r01}
O=gpurun_out
mkdir -p $O
B="python bench.py --steps 3 --warmup 3 --no-cpu --no-hbm-regime"
if [ -z "${ONLY_FULL:-}" ]; then
python bench.py --steps 12 --warmup 3 > $O/bench_${R}_c2.json 2> $O/bench_${R}_c2.err || exit 1
python bench.py --impl reference --steps 4 --warmup 3 > $O/bench_${R}_ref.json 2> $O/bench_${R}_ref.err || exit 1
python bench.py --steps 12 --warmup 3 --no-cpu --no-hbm-regime --workload c3 > $O/bench_${R}_c3.json 2> $O/bench_${R}_c3.err || exit 1
$B > $O/bench_${R}_short.json 2> $O/bench_${R}_short.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_${R}.csv $B > $O/ncu_l.log 2>&1
fi
# 5th resident step (k = 4): each step launches conccalc, pbl, finish once; the 7th+ triples are the
# chunks of the host-buffer part
ncu --set full --clock-control none --import-source on -k 'regex:fpb_(pbl|finish|conccalc)_kernel' \
    --launch-skip 12 --launch-count 3 -f -o $O/full_${R} $B > $O/ncu_f.log 2>&1
ncu -i $O/full_${R}.ncu-rep --page raw --csv > $O/ncu_full_${R}_raw.csv 2>> $O/ncu_f.log
ncu -i $O/full_${R}.ncu-rep --page details > $O/ncu_full_${R}_details.txt 2>> $O/ncu_f.log
tail -2 $O/ncu_f.log
