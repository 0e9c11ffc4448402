python -m pytest tests/test_convection.py -x -q -m gpu 2>&1 | tail -1
for v in "" _r8b3 _k4b4 _r16b1 _k4r16b2; do
  echo "variant libfpb$v: $(FPB_ENGINE_LIB=$PWD/flexpart_b200/libfpb$v.so python tools/convmix_profile.py 2>&1 | tail -1)"
done
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:conv_ --launch-skip 16 -c 8 --csv --log-file gpurun_out/launches_convmix.csv python tools/convmix_profile.py > gpurun_out/ncu_cm.log 2>&1
