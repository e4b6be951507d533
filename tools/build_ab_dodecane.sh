#!/bin/bash
# A/B build of the dodecane_lu kernel TU: tools/build_ab_dodecane.sh <tag> [extra nvcc flags...] -> pelec_b200/lib/libpelerx_b200_<tag>.so
# (LiDryer / drm19 / API objects of the main build are reused; select with PRX_LIB_PATH)
set -u
TAG=$1; shift
cd "$(dirname "$0")/../pelec_b200/csrc"
NVF="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo --extended-lambda -Xcompiler -fPIC -Xcompiler -fvisibility=hidden -Xptxas -v -DPRX_NOINLINE_MATH=1 -DPRX_BLOCK=32 -DPRX_MINBLOCKS=1 -DPRX_RK_BLOCK=32 -DPRX_RK_MINBLOCKS=1 -DPRX_BDF_DEFAULT_VARIANT=32 -DPRX_COOP_MINBLOCKS=${COOPMB:-24} -DPRX_COOP_BLOCK_WARPS=4 -DPRX_VUNROLL_N=1 -DPRX_INLINE_SLOW=1"
mkdir -p ../lib/obj_$TAG
nvcc $NVF -DPRX_MECH_HEADER="\"gen/mech_dodecane_lu.cuh\"" -DPRX_MECH_SUFFIX=dodecane_lu "$@" -c prx_kernels.cu -o ../lib/obj_$TAG/kernels_dodecane_lu.o 2> ../lib/obj_$TAG/kernels_dodecane_lu.log
grep -i "error" -A3 ../lib/obj_$TAG/*.log | head
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../lib/libpelerx_b200_$TAG.so ../lib/obj/api.o ../lib/obj/kernels_LiDryer.o ../lib/obj/kernels_drm19.o ../lib/obj_$TAG/kernels_dodecane_lu.o -Xlinker --version-script=exports.map -lcudart
ls -la ../lib/libpelerx_b200_$TAG.so
grep -h -A3 "bdfc_kernel_dodecane_luILi32ELi4" ../lib/obj_$TAG/kernels_*.log | grep "Used\|stack"
