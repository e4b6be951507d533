#!/bin/bash
# session r02aw: ncu --set full of the final library (384 x 2 blocks)'s default BDF kernel on the TIMED launch size (one 128^3 launch, device-resident path)
set -u
TAG=r02aw
O=pelec_b200/lib/obj
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:bdf_kernel -s 1 -c 1 -f -o gpurun_out/${TAG}_prof_bdf_lidryer_128 \
   python bench.py --steps 1 --warmup 3 --n 128 --no-cpu-baseline --no-ref-gpu --strong none > gpurun_out/${TAG}_ncu1.log 2>&1
tools/ncu_extract.sh gpurun_out/${TAG}_prof_bdf_lidryer_128.ncu-rep $O/kernels_LiDryer.o bdf_kernel_LiDryer
rm -f gpurun_out/${TAG}_prof_bdf_lidryer_128_source.csv.gz
head -40 gpurun_out/${TAG}_prof_bdf_lidryer_128_summary.txt
