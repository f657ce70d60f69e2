#!/bin/bash
mkdir -p gpurun_out
for m in 0 1; do
  AFIGAN_CONV_HALO=$m python tools/profile_one.py 2 1024 1024 200 336 3 && \
  AFIGAN_CONV_HALO=$m ncu --set full --clock-control none --import-source on -k regex:k_conv -s 1 -c 1 -f -o gpurun_out/r2_conv3_halo$m python tools/profile_one.py 2 1024 1024 200 336 3 > gpurun_out/ncu_conv3_halo$m.log 2>&1
  echo "ncu conv3 halo=$m exit $?"
  AFIGAN_CONV_HALO=$m python tools/profile_one.py 2 352 32 104 168 3 && \
  AFIGAN_CONV_HALO=$m ncu --set full --clock-control none --import-source on -k regex:k_conv -s 1 -c 1 -f -o gpurun_out/r2_growth_halo$m python tools/profile_one.py 2 352 32 104 168 3 > gpurun_out/ncu_growth_halo$m.log 2>&1
  echo "ncu growth halo=$m exit $?"
done
