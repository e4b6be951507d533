#!/bin/bash
# A/B builds of the LiDryer kernel TU (slim: gen-1 BDF, cooperative BDF, RK64):
#   tools/build_ab.sh <tag> <header: c|nc> [extra nvcc flags...]  ->  pelec_b200/lib/libpelerx_b200_<tag>.so (select with PRX_LIB_PATH)
# header c = literal pool (gen/mech_LiDryer.cuh), nc = inline literals (gen/mech_LiDryer_nc.cuh, made with PRX_GEN_CONSTIFY=0)
set -u
TAG=$1; HDR=$2; shift 2
cd "$(dirname "$0")/../pelec_b200/csrc"
H=gen/mech_LiDryer.cuh; [ "$HDR" = nc ] && H=gen/mech_LiDryer_nc.cuh
NVF="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo --extended-lambda -Xcompiler -fPIC -Xcompiler -fvisibility=hidden -Xptxas -v -DPRX_COOP_MINBLOCKS=24 -DPRX_VUNROLL_N=32 -DPRX_BDF_DEFAULT_VARIANT=0 -DPRX_INLINE_SLOW=1"
mkdir -p ../lib/obj_$TAG
nvcc $NVF -DPRX_MECH_HEADER="\"$H\"" -DPRX_MECH_SUFFIX=LiDryer -DPRX_BLOCK=${LB:-224} -DPRX_MINBLOCKS=${LM:-3} -DPRX_RK_BLOCK=${RB:-256} -DPRX_RK_MINBLOCKS=${RM:-2} -DPRX_COOP_BLOCK_WARPS=1 "$@" -c prx_kernels.cu -o ../lib/obj_$TAG/kernels_LiDryer.o 2> ../lib/obj_$TAG/kernels_LiDryer.log
grep -i "error" -A3 ../lib/obj_$TAG/*.log | head
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../lib/libpelerx_b200_$TAG.so ../lib/obj/api.o ../lib/obj_$TAG/kernels_LiDryer.o \
   $(ls ../lib/obj/kernels_*.o | grep -v kernels_LiDryer.o) -Xlinker --version-script=exports.map -lcudart
ls -la ../lib/libpelerx_b200_$TAG.so
grep -h -A2 "18bdf_kernel\|rk64_kernel" ../lib/obj_$TAG/kernels_*.log | grep "Used\|stack"
