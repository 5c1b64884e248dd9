#!/bin/bash
# parity + timing of the kernel paths on one GPU (run under gpurun)
mkdir -p gpurun_out
timeout 300 python tools/ab_check.py > gpurun_out/ab_check.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/ab_tests.txt
{
echo "== default (ws level 0 + fused)"; timeout 120 python tools/time_clip.py --fps 30 --steps 10
echo "== fused only"; FVVDP_B200_PATH=fused timeout 120 python tools/time_clip.py --fps 30 --steps 10
echo "== ws all levels"; FVVDP_B200_WS_LEVELS=7 timeout 120 python tools/time_clip.py --fps 30 --steps 10
echo "== 1080p"; timeout 120 python tools/time_clip.py --fps 30 --size 1920x1080 --display standard_fhd --steps 10
echo "== 60 fps"; timeout 120 python tools/time_clip.py --fps 60 --steps 5
echo "== 24 fps"; timeout 120 python tools/time_clip.py --fps 24 --steps 5
echo "== 120 fps"; timeout 120 python tools/time_clip.py --fps 120 --steps 3
echo "== fov default"; timeout 120 python tools/time_clip.py --fps 30 --display standard_hdr_pq --foveated --steps 5
echo "== fov fused"; FVVDP_B200_PATH=fused timeout 120 python tools/time_clip.py --fps 30 --display standard_hdr_pq --foveated --steps 5
} > gpurun_out/ab_times.txt 2>&1
tail -4 gpurun_out/ab_check.txt; cat gpurun_out/ab_tests.txt gpurun_out/ab_times.txt
