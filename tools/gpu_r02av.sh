#!/bin/bash
# session r02av: run-time barrier modes re-measured with the 384 x 2 block shape (PRX_BLOCK_WARPS: 0 both phase barriers, 2 only the one
# before the RHS, 3 only the exit test, 1 none)
mkdir -p gpurun_out
for m in 0 2 3 1; do
  PRX_BLOCK_WARPS=$m python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-ref-gpu --strong none > gpurun_out/r02av_bench_lidryer_bw$m.json 2> gpurun_out/r02av_bw$m.err
  python - <<P
import json
try:
    d=json.loads(open("gpurun_out/r02av_bench_lidryer_bw$m.json").read().strip().splitlines()[-1])
    print("PRX_BLOCK_WARPS=$m: %.2f ms, failed %s"%(d["ms_per_step"], d["failed_cells"]))
except Exception as e:
    print("bw$m FAILED", e); print(open("gpurun_out/r02av_bw$m.err").read()[-300:])
P
done
