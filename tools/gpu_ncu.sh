#!/bin/bash
# ncu --set full capture of one kernel: tools/gpu_ncu.sh <tag> <kernel-regex> <bench args...>
set -u
TAG=$1; KREG=$2; shift 2
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$KREG -s 1 -c 1 -o gpurun_out/prof_$TAG -f \
   python bench.py --steps 1 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/ncu_$TAG.log 2>&1
tail -3 gpurun_out/ncu_$TAG.log
