#!/bin/bash
# time variant builds of the library (tools/build_variants.py) on the 4K bench clip; args: variant names
mkdir -p gpurun_out
: > gpurun_out/var_times.txt
for v in "$@"; do
  L=$PWD/fovvideovdp_b200/_lib/variants/$v/libfvvdp_b200.so
  echo "== $v" >> gpurun_out/var_times.txt
  FVVDP_B200_LIB=$L FVVDP_B200_WS_LEVELS=${WSL:-1} timeout 120 python tools/time_clip.py --fps 30 --steps 10 >> gpurun_out/var_times.txt 2>&1
  FVVDP_B200_LIB=$L timeout 120 python tools/time_clip.py --fps 30 --steps 10 >> gpurun_out/var_times.txt 2>&1
done
cat gpurun_out/var_times.txt
