#!/bin/bash
# one ncu --set full capture of the level-0 / level-1 band kernels of the bench clip (run under gpurun; env WSL = ws levels)
mkdir -p gpurun_out
FVVDP_B200_WS_LEVELS=${WSL:-7} timeout 900 ncu --set full --clock-control none --import-source on -k regex:band_ -c 2 -f -o gpurun_out/${OUT:-prof} python tools/time_clip.py --fps 30 --steps 1 > gpurun_out/ncu_log.txt 2>&1
tail -3 gpurun_out/ncu_log.txt
