#!/bin/bash
# A/B builds of libls3d_b200.so with extra -D flags:  scripts/ab_build.sh <tag> [-DNAME=VALUE ...]  ->  livescan3d_b200/csrc/build/ab/libls3d_<tag>.so
set -e
cd "$(dirname "$0")/../livescan3d_b200/csrc"
tag=$1; shift
out=build/ab; mkdir -p $out/obj_$tag
for f in runtime frame icp preprocess formats; do
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false -Xcompiler -fPIC -Xptxas -v -cudart static "$@" -c $f.cu -o $out/obj_$tag/$f.o 2> $out/obj_$tag/$f.ptxas.log &
done
wait
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -cudart static -o $out/libls3d_$tag.so $out/obj_$tag/runtime.o $out/obj_$tag/frame.o $out/obj_$tag/icp.o $out/obj_$tag/preprocess.o $out/obj_$tag/formats.o -ldl
grep -A2 -E "k_organized_count|k_icp_match_packet" $out/obj_$tag/*.ptxas.log | grep -E "Used|spill" | head -6
