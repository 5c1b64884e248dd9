#!/bin/bash
# robustness checks on one GPU: random configurations (default kernel paths, then the warp-specialised kernel on every level)
# against the general kernels, and compute-sanitizer memcheck (RACE=1: also racecheck) over every kernel family on small clips
mkdir -p gpurun_out
{
echo "== fuzz, default paths vs general kernels"; timeout 600 python tools/fuzz_paths.py 1 150 2>&1 | tail -6
echo "== fuzz, warp-specialised kernel on every level vs general kernels"; FVVDP_B200_WS_LEVELS=7 timeout 600 python tools/fuzz_paths.py 2 150 2>&1 | tail -6
} > gpurun_out/r02_fuzz_paths.txt 2>&1
{
echo "== memcheck"; FVVDP_B200_WS_LEVELS=7 timeout 900 compute-sanitizer --tool memcheck python tools/sanitize_small.py 2>&1 | tail -34
if [ -n "$RACE" ]; then echo "== racecheck"; FVVDP_B200_WS_LEVELS=7 timeout 1200 compute-sanitizer --tool racecheck python tools/sanitize_small.py 2>&1 | tail -14; fi
} > gpurun_out/r02_compute_sanitizer_memcheck.txt 2>&1
cat gpurun_out/r02_fuzz_paths.txt gpurun_out/r02_compute_sanitizer_memcheck.txt
