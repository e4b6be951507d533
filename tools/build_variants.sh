#!/bin/bash
# A/B builds of the LiDryer / drm19 kernel translation units (block size, min blocks per SM) into
# pelec_b200/lib/libpelerx_b200_<tag>.so, selected at run time with PRX_LIB_PATH.  Scratch tooling for GPU sessions.
set -u
cd "$(dirname "$0")/../pelec_b200/csrc"
NVF="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo --extended-lambda -Xcompiler -fPIC -Xcompiler -fvisibility=hidden -Xptxas -v -DPRX_NOINLINE_MATH=1"
build() { # tag mech block minblocks variant
  local tag=$1 m=$2 b=$3 mb=$4 v=$5
  mkdir -p ../lib/obj_$tag
  nvcc $NVF -DPRX_MECH_HEADER="\"gen/mech_$m.cuh\"" -DPRX_MECH_SUFFIX=$m -DPRX_BLOCK=$b -DPRX_MINBLOCKS=$mb -DPRX_BDF_DEFAULT_VARIANT=$v \
     -DPRX_COOP_MINBLOCKS=24 -c prx_kernels.cu -o ../lib/obj_$tag/kernels_$m.o 2> ../lib/obj_$tag/kernels_$m.log
}
for spec in "b256m2 256 2" "b128m4 128 4" "b128m6 128 6" "b64m8 64 8" "b256m1 256 1"; do
  set -- $spec
  build $1 LiDryer $2 $3 0 &
done
build d256m1 drm19 256 1 0 &
build d128m2 drm19 128 2 0 &
build d128m4 drm19 128 4 0 &
wait
for tag in b256m2 b128m4 b128m6 b64m8 b256m1; do
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../lib/libpelerx_b200_$tag.so ../lib/obj/api.o ../lib/obj_$tag/kernels_LiDryer.o \
     ../lib/obj_d256m1/kernels_drm19.o ../lib/obj/kernels_dodecane_lu.o -Xlinker --version-script=exports.map -lcudart
done
for tag in d128m2 d128m4; do
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../lib/libpelerx_b200_$tag.so ../lib/obj/api.o ../lib/obj_b256m2/kernels_LiDryer.o \
     ../lib/obj_$tag/kernels_drm19.o ../lib/obj/kernels_dodecane_lu.o -Xlinker --version-script=exports.map -lcudart
done
ls -la ../lib/*.so
grep -h -A2 "18bdf_kernel" ../lib/obj_*/kernels_*.log | grep "Used\|stack"
