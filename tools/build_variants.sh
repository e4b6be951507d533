#!/bin/bash
# A/B build of the LiDryer / drm19 kernel translation units with extra flags:
#   tools/build_variants.sh <tag> [extra nvcc flags...]   ->  pelec_b200/lib/libpelerx_b200_<tag>.so  (select with PRX_LIB_PATH)
# Scratch tooling for GPU sessions; the dodecane_lu and API objects of the main build are reused.
set -u
TAG=$1; shift
cd "$(dirname "$0")/../pelec_b200/csrc"
NVF="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo --extended-lambda -Xcompiler -fPIC -Xcompiler -fvisibility=hidden -Xptxas -v -DPRX_NOINLINE_MATH=1 -DPRX_COOP_MINBLOCKS=24 -DPRX_VUNROLL_N=32 -DPRX_BDF_DEFAULT_VARIANT=0"
mkdir -p ../lib/obj_$TAG
nvcc $NVF -DPRX_MECH_HEADER="\"gen/mech_LiDryer.cuh\"" -DPRX_MECH_SUFFIX=LiDryer -DPRX_BLOCK=${LB:-256} -DPRX_MINBLOCKS=${LM:-2} "$@" -c prx_kernels.cu -o ../lib/obj_$TAG/kernels_LiDryer.o 2> ../lib/obj_$TAG/kernels_LiDryer.log &
nvcc $NVF -DPRX_MECH_HEADER="\"gen/mech_drm19.cuh\"" -DPRX_MECH_SUFFIX=drm19 -DPRX_BLOCK=${DB:-128} -DPRX_MINBLOCKS=${DM:-4} "$@" -c prx_kernels.cu -o ../lib/obj_$TAG/kernels_drm19.o 2> ../lib/obj_$TAG/kernels_drm19.log &
wait
grep -i "error" -A3 ../lib/obj_$TAG/*.log | head
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../lib/libpelerx_b200_$TAG.so ../lib/obj/api.o ../lib/obj_$TAG/kernels_LiDryer.o \
   ../lib/obj_$TAG/kernels_drm19.o ../lib/obj/kernels_dodecane_lu.o -Xlinker --version-script=exports.map -lcudart
ls -la ../lib/libpelerx_b200_$TAG.so
grep -h -A2 "18bdf_kernel" ../lib/obj_$TAG/kernels_*.log | grep "Used\|stack"
