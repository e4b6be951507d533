#!/bin/bash
# session r02ar: gen-1 BDF block shapes that phase-align ALL warps of an SM (one block per SM): 672 x 1 (80 registers, 21 warps),
# 448 x 1 (128 registers, 14 warps), and 352 x 2 (80 registers, 22 warps) against the shipped 224 x 3
mkdir -p gpurun_out
run() { # tag lib
  PRX_LIB_PATH=$2 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-ref-gpu --strong none > gpurun_out/r02ar_bench_lidryer_$1.json 2> gpurun_out/r02ar_$1.err
  python - <<P
import json
try:
    d=json.loads(open("gpurun_out/r02ar_bench_lidryer_$1.json").read().strip().splitlines()[-1])
    print("$1: %.2f ms, failed %s"%(d["ms_per_step"], d["failed_cells"]))
except Exception as e:
    print("$1 FAILED", e); print(open("gpurun_out/r02ar_$1.err").read()[-300:])
P
}
for rep in 1 2; do
run base_$rep pelec_b200/lib/libpelerx_b200.so
for v in b352 b384; do run ${v}_$rep pelec_b200/lib/libpelerx_b200_$v.so; done
done
