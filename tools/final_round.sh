#!/bin/bash
# tools/final_round.sh -- on the GPU box: the whole GPU suite, the un-profiled C2 bench line, then the convmix launch
# list and one --set full capture of its three column kernels (tools/convmix_profile.py)
O=gpurun_out
python -m pytest tests -m gpu -x -q > $O/full_gpu.log 2>&1; tail -3 $O/full_gpu.log
python bench.py --steps 12 --warmup 3 > $O/bench_r02_c2_final.json 2> $O/bench_r02_c2_final.err || tail -5 $O/bench_r02_c2_final.err
python tools/convmix_profile.py > $O/convmix_final.log 2>&1; tail -1 $O/convmix_final.log
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:conv_ --launch-skip 16 -c 8 --csv --log-file $O/launches_convmix.csv python tools/convmix_profile.py > $O/ncu_cm.log 2>&1
ncu --set full --clock-control none --import-source on -k "regex:conv_(column_kernel|mix|assembly)" --launch-skip 6 --launch-count 3 -f -o $O/ncu_convmix_final python tools/convmix_profile.py > $O/ncu_cm2.log 2>&1
ncu -i $O/ncu_convmix_final.ncu-rep --page details > $O/ncu_convmix_final_details.txt 2>> $O/ncu_cm2.log
