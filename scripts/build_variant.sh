#!/bin/bash
# Build an experimental variant of the library next to the product one:
#   scripts/build_variant.sh NAME -DCMAX_KNN_TILE_W=32 ...   ->  _variants/NAME/libcmax_b200.so
# (a GPU run swaps it in with: cp _variants/NAME/libcmax_b200.so motionpriorcmax_b200/_lib/)
set -e
name=$1; shift
root=$(cd "$(dirname "$0")/.." && pwd)
out=$root/_variants/$name
mkdir -p $out
csrc=$root/motionpriorcmax_b200/csrc
flags="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false -Xcompiler -fPIC -Xcompiler -O2"
for s in api lut_stage event_stage tile_stage image_stage voxel_stage flow_stage; do
  nvcc $flags "$@" -c $csrc/$s.cu -o $out/$s.o &
done
g++ -O3 -msse4.1 -std=c++17 -fPIC -fopenmp -c $csrc/host_pack.cpp -o $out/host_pack.o
wait
nvcc -shared -o $out/libcmax_b200.so $out/*.o -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -cudart shared -lgomp
rm $out/*.o
echo $out/libcmax_b200.so
