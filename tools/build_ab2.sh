#!/bin/bash
# A/B build of ONE mechanism's kernel TU with other block shapes / flags (the other TUs and api.o come from the main build):
#   MECH=drm19 LB=256 LM=2 tools/build_ab2.sh <tag> [extra nvcc flags...]  ->  pelec_b200/lib/libpelerx_b200_<tag>.so (PRX_LIB_PATH)
set -u
TAG=$1; shift
MECH=${MECH:-LiDryer}
cd "$(dirname "$0")/../pelec_b200/csrc"
NVF="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo --extended-lambda -Xcompiler -fPIC -Xcompiler -fvisibility=hidden -Xptxas -v -DPRX_NOINLINE_MATH=1 -DPRX_COOP_MINBLOCKS=24 -DPRX_VUNROLL_N=32 -DPRX_BDF_DEFAULT_VARIANT=0 -DPRX_INLINE_SLOW=1"
mkdir -p ../lib/obj_$TAG
nvcc $NVF -DPRX_MECH_HEADER="\"gen/mech_$MECH.cuh\"" -DPRX_MECH_SUFFIX=$MECH -DPRX_BLOCK=${LB:-224} -DPRX_MINBLOCKS=${LM:-3} -DPRX_RK_BLOCK=${RB:-256} -DPRX_RK_MINBLOCKS=${RM:-2} -DPRX_COOP_BLOCK_WARPS=1 "$@" -c prx_kernels.cu -o ../lib/obj_$TAG/kernels_$MECH.o 2> ../lib/obj_$TAG/kernels_$MECH.log
grep -i "error" -A3 ../lib/obj_$TAG/*.log | head
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../lib/libpelerx_b200_$TAG.so ../lib/obj/api.o ../lib/obj_$TAG/kernels_$MECH.o \
   $(ls ../lib/obj/kernels_*.o | grep -v kernels_$MECH.o) -Xlinker --version-script=exports.map -lcudart
grep -h -A2 "18bdf_kernel" ../lib/obj_$TAG/kernels_*.log | grep "Used\|stack"
