#!/bin/bash
# tools/profile_round.sh ROUND -- run on the GPU box (gpurun -- 'bash tools/profile_round.sh r02'):
# the un-profiled bench lines first, then (same commands, after they exited 0) the ncu launch lists
# and one --set full capture of the hot-path kernels of the C2, C3 and C5 workloads.  Everything lands
# in gpurun_out/; tools/summarize_profiles.py turns it into profiles/.
set -u
R=${1:-r02}
O=gpurun_out
mkdir -p $O
B="python bench.py --steps 3 --warmup 3 --no-cpu --no-hbm-regime --no-c5"
B3="$B --workload c3"
B5="python bench.py --workload c5 --steps 3 --warmup 3"
python bench.py --steps 12 --warmup 3 > $O/bench_${R}_c2.json 2> $O/bench_${R}_c2.err || exit 1
python bench.py --impl reference --steps 4 --warmup 3 > $O/bench_${R}_ref.json 2> $O/bench_${R}_ref.err || exit 1
python bench.py --steps 12 --warmup 3 --no-cpu --no-hbm-regime --no-c5 --workload c3 > $O/bench_${R}_c3.json 2> $O/bench_${R}_c3.err || exit 1
python bench.py --workload c5 --steps 12 --warmup 3 > $O/bench_${R}_c5.json 2> $O/bench_${R}_c5.err || exit 1
$B > $O/bench_${R}_short.json 2> $O/bench_${R}_short.err || exit 1
# ---- C2: launch list + full capture (5th resident step: conccalc, pbl, finish)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_${R}.csv $B > $O/ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k 'regex:fpb_(pbl|finish|conccalc)_kernel' \
    --launch-skip 12 --launch-count 3 -f -o $O/full_${R} $B > $O/ncu_f.log 2>&1
ncu -i $O/full_${R}.ncu-rep --page raw --csv > $O/ncu_full_${R}_raw.csv 2>> $O/ncu_f.log
ncu -i $O/full_${R}.ncu-rep --page details > $O/ncu_full_${R}_details.txt 2>> $O/ncu_f.log
# ---- C3 (CBL + deposition + nested output): the full-feature instantiation fpb_pbl_kernel<1,1,0>
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_${R}_c3.csv $B3 > $O/ncu_l3.log 2>&1
ncu --set full --clock-control none --import-source on -k 'regex:fpb_(pbl|finish|conccalc|wetdepo)_kernel' \
    --launch-skip 15 --launch-count 4 -f -o $O/full_${R}_c3 $B3 > $O/ncu_f3.log 2>&1
ncu -i $O/full_${R}_c3.ncu-rep --page raw --csv > $O/ncu_full_${R}_c3_raw.csv 2>> $O/ncu_f3.log
ncu -i $O/full_${R}_c3.ncu-rep --page details > $O/ncu_full_${R}_c3_details.txt 2>> $O/ncu_f3.log
# ---- C5 (100 M domain-filling particles, method 0): the HBM-gather regime
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $O/launches_${R}_c5.csv $B5 > $O/ncu_l5.log 2>&1
ncu --set full --clock-control none --import-source on -k 'regex:fpb_(pbl|finish|conccalc)_kernel' \
    --launch-skip 9 --launch-count 3 -f -o $O/full_${R}_c5 $B5 > $O/ncu_f5.log 2>&1
ncu -i $O/full_${R}_c5.ncu-rep --page raw --csv > $O/ncu_full_${R}_c5_raw.csv 2>> $O/ncu_f5.log
ncu -i $O/full_${R}_c5.ncu-rep --page details > $O/ncu_full_${R}_c5_details.txt 2>> $O/ncu_f5.log
# ---- calcpar + verttransform_ecmwf on the device: launch list of the met_* kernels
python tools/metproc_profile.py > $O/metproc_${R}.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k 'regex:(met_|pack_group|pair_)' \
    -c 60 --csv --log-file $O/launches_${R}_metproc.csv python tools/metproc_profile.py > $O/ncu_lm.log 2>&1
rm -f $O/full_${R}*.ncu-rep   # (the exports above are what is tracked; the reports exceed the 64 MiB merge limit)
tail -2 $O/ncu_f.log $O/ncu_f3.log $O/ncu_f5.log
