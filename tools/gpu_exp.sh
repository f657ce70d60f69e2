#!/bin/bash
for cfg in "1 3" "1 2"; do set -- $cfg
  AFIGAN_HALO_DBG=1 AFIGAN_CONV_HALO=$1 AFIGAN_HALO_SA=$2 python tools/profile_one.py 2 1024 1024 200 336 2 2>&1 | tail -3
  AFIGAN_HALO_DBG=1 AFIGAN_CONV_HALO=$1 AFIGAN_HALO_SA=$2 python tools/profile_one.py 2 352 32 104 168 2 2>&1 | tail -3
  AFIGAN_CONV_HALO=$1 AFIGAN_HALO_SA=$2 python tools/profile_one.py 2 1024 1024 200 336 10
  AFIGAN_CONV_HALO=$1 AFIGAN_HALO_SA=$2 python tools/profile_one.py 2 352 32 104 168 10
done
