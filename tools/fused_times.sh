for v in "" np; do
  if [ -n "$v" ]; then export FVVDP_B200_LIB=$PWD/fovvideovdp_b200/_lib/variants/$v/libfvvdp_b200.so; fi
  echo "== variant '$v' fused only"; FVVDP_B200_PATH=fused timeout 120 python tools/time_clip.py --fps 30 --steps 10
  echo "== variant '$v' default"; timeout 120 python tools/time_clip.py --fps 30 --steps 10
  echo "== variant '$v' 60fps"; timeout 120 python tools/time_clip.py --fps 60 --steps 5
done
