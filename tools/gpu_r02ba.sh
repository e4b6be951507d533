#!/bin/bash
# session r02ba: slab count of the host staging pipeline re-measured with the final (faster) kernel
mkdir -p gpurun_out
for s in 8 12 14 6; do
  PRX_HOST_SLABS=$s python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-ref-gpu --strong none > gpurun_out/r02ba_bench_slabs$s.json 2> gpurun_out/r02ba_$s.err
  python - <<P
import json
try:
    d=json.loads(open("gpurun_out/r02ba_bench_slabs$s.json").read().strip().splitlines()[-1])
    print("slabs $s: device %.2f ms, e2e %.2f ms (%.2f M cells/s)"%(d["ms_per_step"], d["e2e"]["ms_per_step"], d["e2e"]["value"]/1e6))
except Exception as e:
    print("slabs $s FAILED", e)
P
done
