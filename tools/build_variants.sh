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
     -DPRX_COOP_MINBLOCKS=24 $EXTRA -c prx_kernels.cu -o ../lib/obj_$tag/kernels_$m.o 2> ../lib/obj_$tag/kernels_$m.log
}
EXTRA=""
build vu1 LiDryer 256 2 0 &
build dvu1 drm19 128 4 0 &
EXTRA="-DPRX_VUNROLL_N=32"
build vu32 LiDryer 256 2 0 &
build dvu32 drm19 128 4 0 &
wait
for t in 1 32; do
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../lib/libpelerx_b200_vu$t.so ../lib/obj/api.o ../lib/obj_vu$t/kernels_LiDryer.o \
     ../lib/obj_dvu$t/kernels_drm19.o ../lib/obj/kernels_dodecane_lu.o -Xlinker --version-script=exports.map -lcudart
done
ls -la ../lib/*.so
for t in vu1 vu32 dvu1 dvu32; do grep -h -A2 "18bdf_kernel" ../lib/obj_$t/kernels_*.log | grep "Used\|stack"; done
