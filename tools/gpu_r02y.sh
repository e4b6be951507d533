#!/bin/bash
# session r02y: state prefetch behind the RHS (PRX_PREFETCH_STATE A/B), pointwise kernels after the occupancy / unroll tweak
mkdir -p gpurun_out
for v in base pf1 pf3 pf5; do
  lib=pelec_b200/lib/libpelerx_b200_$v.so; [ $v = base ] && lib=pelec_b200/lib/libpelerx_b200.so
  for rep in 1 2; do
  PRX_LIB_PATH=$lib python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-ref-gpu --strong none > gpurun_out/r02y_bench_lidryer_${v}_$rep.json 2> gpurun_out/r02y_$v.err
  python - <<P
import json
try:
    d=json.loads(open("gpurun_out/r02y_bench_lidryer_${v}_$rep.json").read().strip().splitlines()[-1])
    print("$v run $rep: %.2f ms, failed %s"%(d["ms_per_step"], d["failed_cells"]))
except Exception as e:
    print("$v FAILED", e); print(open("gpurun_out/r02y_$v.err").read()[-300:])
P
  done
done
for k in ctoprim transport; do python tools/bench_eos.py --mech LiDryer --n 192 --kernel $k | tee gpurun_out/r02y_eos_${k}_LiDryer_192.json | cut -c1-160; done
