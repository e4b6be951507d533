#!/bin/bash
# session r02ax: out-of-line Jacobian (1) / RHS (2) / both (3) re-measured with the 384 x 2 block shape (the r02aw capture: the
# kernel is sensitive to the amount of distinct code its phase groups walk)
mkdir -p gpurun_out
run() { # tag lib
  PRX_LIB_PATH=$2 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-ref-gpu --strong none > gpurun_out/r02ax_bench_lidryer_$1.json 2> gpurun_out/r02ax_$1.err
  python - <<P
import json
try:
    d=json.loads(open("gpurun_out/r02ax_bench_lidryer_$1.json").read().strip().splitlines()[-1])
    print("$1: %.2f ms, failed %s"%(d["ms_per_step"], d["failed_cells"]))
except Exception as e:
    print("$1 FAILED", e); print(open("gpurun_out/r02ax_$1.err").read()[-300:])
P
}
L=pelec_b200/lib/libpelerx_b200
for rep in 1 2; do
run base_$rep $L.so
for v in o1 o2 o3; do run ${v}_$rep ${L}_$v.so; done
done
