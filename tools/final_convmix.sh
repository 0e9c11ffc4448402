#!/bin/bash
# tools/final_convmix.sh -- on the GPU box: the un-profiled C2 bench line, then fpb_convmix alone (tools/convmix_profile.py):
# launch list and one --set full capture of its column kernels
O=gpurun_out
python bench.py --steps 12 --warmup 3 > $O/bench_r02_c2_final.json 2> $O/bench_r02_c2_final.err || tail -5 $O/bench_r02_c2_final.err
python tools/convmix_profile.py > $O/convmix_final.log 2>&1; tail -1 $O/convmix_final.log
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:conv_ --launch-skip 20 -c 10 --csv --log-file $O/launches_convmix.csv python tools/convmix_profile.py > $O/ncu_cm.log 2>&1
ncu --set full --clock-control none --import-source on -k "regex:conv_(pre|column_kernel|mix|assembly)" --launch-skip 8 --launch-count 4 -f -o $O/ncu_convmix_final python tools/convmix_profile.py > $O/ncu_cm2.log 2>&1
