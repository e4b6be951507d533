#!/bin/bash
# session r02as: gen-1 BDF block shapes, continued: LiDryer 384 x 2 and 256 x 3 (both 24 warps, 80 registers) vs 224 x 3; drm19 256 x 2
# (128 registers, 16 warps) and 384 x 1 (168 registers, 12 warps in one block) vs 192 x 2
mkdir -p gpurun_out
run() { # tag lib extra-args
  PRX_LIB_PATH=$2 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-ref-gpu --strong none $3 > gpurun_out/r02as_bench_$1.json 2> gpurun_out/r02as_$1.err
  python - <<P
import json
try:
    d=json.loads(open("gpurun_out/r02as_bench_$1.json").read().strip().splitlines()[-1])
    print("$1: %.2f ms, failed %s"%(d["ms_per_step"], d["failed_cells"]))
except Exception as e:
    print("$1 FAILED", e); print(open("gpurun_out/r02as_$1.err").read()[-300:])
P
}
L=pelec_b200/lib/libpelerx_b200
for rep in 1 2; do
run lidryer_base_$rep $L.so ""
run lidryer_b384_$rep ${L}_b384.so ""
run lidryer_b256_$rep ${L}_b256.so ""
run drm19_base_$rep $L.so "--mech drm19 --n 64"
run drm19_d256_$rep ${L}_d256.so "--mech drm19 --n 64"
run drm19_d384_$rep ${L}_d384.so "--mech drm19 --n 64"
done
