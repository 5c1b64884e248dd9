#!/bin/bash
# build an experiment variant of the library: tools/build_variant.sh <name> [extra nvcc flags...] -> gpurun_out/variants/lib_<name>.so
set -e
name=$1; shift
cd /root/repo/fovvideovdp_b200/csrc
out=/root/repo/build/variants; mkdir -p $out/obj_$name
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC $@"
nvcc $FLAGS -c fvvdp_b200.cu -o $out/obj_$name/fvvdp_b200.o
nvcc $FLAGS -c fvvdp_fused_dispatch.cu -o $out/obj_$name/d.o &
for k in 0 2 3; do for v in 0 1 2; do nvcc $FLAGS -DFUSED_KIND=$k -DFUSED_VIDEO=$v -c fvvdp_fused_inst.cu -o $out/obj_$name/f_${k}_${v}.o & done; done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -Xcompiler -fPIC -o $out/lib_$name.so $out/obj_$name/*.o
rm -rf $out/obj_$name
echo built $out/lib_$name.so
