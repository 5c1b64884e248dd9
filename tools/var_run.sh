#!/bin/bash
# time variant builds of the library (tools/build_variants.py) on the 4K bench clip; args: variant names
# every run starts with the fused-kernel path of the main build as the box's yardstick; WSLS = list of ws level counts to time
mkdir -p gpurun_out
: > gpurun_out/var_times.txt
echo "== fused (yardstick)" >> gpurun_out/var_times.txt
FVVDP_B200_PATH=fused timeout 120 python tools/time_clip.py --fps 30 --steps 10 >> gpurun_out/var_times.txt 2>&1
for v in "$@"; do
  L=$PWD/fovvideovdp_b200/_lib/variants/$v/libfvvdp_b200.so
  for wl in ${WSLS:-1 7}; do
    echo "== $v, warp-specialised kernel on levels < $wl" >> gpurun_out/var_times.txt
    FVVDP_B200_LIB=$L FVVDP_B200_WS_LEVELS=$wl timeout 120 python tools/time_clip.py --fps 30 --steps 10 >> gpurun_out/var_times.txt 2>&1
  done
done
cat gpurun_out/var_times.txt
