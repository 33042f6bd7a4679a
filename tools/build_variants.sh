#!/bin/bash
# build variants of one translation unit with different -D flags into variants/<name>.so
#   usage: build_variants.sh <extend|shade> name "flags" [name "flags" ...]
mkdir -p /root/repo/variants; cd /root/repo/nori-ray-tracer_b200/csrc
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -Xcompiler -fPIC -I../../include"
tu=$1; shift
while [ $# -gt 1 ]; do
  name=$1; defs=$2; shift 2
  if [ $tu = extend ]; then
    ( nvcc $FLAGS $defs -c -o /tmp/v_$name.o wave_extend.cu 2>/dev/null && \
      nvcc -gencode arch=compute_100a,code=sm_100a -shared -o /root/repo/variants/$name.so obj/nori_gpu.o /tmp/v_$name.o obj/wave_shade0.o obj/wave_shade1.o obj/wave_shade2.o obj/mega.o obj/wave_drain.o obj/gpu_bvh.o obj/host_bvh.o obj/wave_extend_p.o obj/wave_shade_p0.o obj/wave_shade_p1.o obj/wave_shade_p2.o obj/wave_drain_p.o && echo built $name ) &
  else
    ( nvcc $FLAGS $defs -DNORI_SHADE_MODE=1 -DNORI_SHADE_TEMPLATED=1 -DNORI_DYN_INLINE=1 -c -o /tmp/v_$name.o wave_shade.cu 2>/dev/null && \
      nvcc -gencode arch=compute_100a,code=sm_100a -shared -o /root/repo/variants/$name.so obj/nori_gpu.o obj/wave_extend.o obj/wave_shade0.o /tmp/v_$name.o obj/wave_shade2.o obj/mega.o obj/wave_drain.o obj/gpu_bvh.o obj/host_bvh.o obj/wave_extend_p.o obj/wave_shade_p0.o obj/wave_shade_p1.o obj/wave_shade_p2.o obj/wave_drain_p.o && echo built $name ) &
  fi
done
wait
