#!/bin/bash
# session r02at: more resident warps at fewer registers: 416 x 2 (26 warps, 72 registers), 448 x 2 (28, 72), 512 x 2 (32, 64) vs 384 x 2 (24, 80)
mkdir -p gpurun_out
run() { # tag lib
  PRX_LIB_PATH=$2 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-ref-gpu --strong none > gpurun_out/r02at_bench_lidryer_$1.json 2> gpurun_out/r02at_$1.err
  python - <<P
import json
try:
    d=json.loads(open("gpurun_out/r02at_bench_lidryer_$1.json").read().strip().splitlines()[-1])
    print("$1: %.2f ms, failed %s"%(d["ms_per_step"], d["failed_cells"]))
except Exception as e:
    print("$1 FAILED", e); print(open("gpurun_out/r02at_$1.err").read()[-300:])
P
}
L=pelec_b200/lib/libpelerx_b200
for rep in 1 2; do
for v in b384 b416 b448x2 b512; do run ${v}_$rep ${L}_$v.so; done
done
