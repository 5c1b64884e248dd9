#!/bin/bash
# A/B timing of the kernel paths on one GPU (run under gpurun): parity checks first, then 4K / 1080p clips per path
mkdir -p gpurun_out
timeout 300 python tools/ab_check.py > gpurun_out/ab_check.txt 2>&1
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/ab_tests.txt
{
for lv in 1 2 3 7; do
  echo "== WS_LEVELS=$lv"
  FVVDP_B200_WS_LEVELS=$lv timeout 120 python tools/time_clip.py --fps 30 --steps 10
done
echo "== fused"; FVVDP_B200_PATH=fused timeout 120 python tools/time_clip.py --fps 30 --steps 10
echo "== 1080p ws"; timeout 120 python tools/time_clip.py --fps 30 --size 1920x1080 --display standard_fhd --steps 10
echo "== 1080p fused"; FVVDP_B200_PATH=fused timeout 120 python tools/time_clip.py --fps 30 --size 1920x1080 --display standard_fhd --steps 10
echo "== fov ws"; timeout 120 python tools/time_clip.py --fps 30 --display standard_hdr_pq --foveated --steps 5
echo "== fov fused"; FVVDP_B200_PATH=fused timeout 120 python tools/time_clip.py --fps 30 --display standard_hdr_pq --foveated --steps 5
} > gpurun_out/ab_times.txt 2>&1
cat gpurun_out/ab_check.txt gpurun_out/ab_tests.txt gpurun_out/ab_times.txt
if [ -n "$NCU" ]; then
timeout 600 ncu --set full --clock-control none --import-source on -k regex:band_ws -c 2 -f -o gpurun_out/ws_prof python tools/time_clip.py --fps 30 --steps 1 > gpurun_out/ncu_log.txt 2>&1
fi
