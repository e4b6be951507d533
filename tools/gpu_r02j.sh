#!/bin/bash
# Round 2, session J: RK64 block-shape A/B (r1 192x3, r2 128x4, r3 384x1, r4 160x3 vs default 256x2), LiDryer + flame states
set -u
TAG=r02j
mkdir -p gpurun_out
for v in r1 r2 r3 r4; do
  export PRX_LIB_PATH=$PWD/pelec_b200/lib/libpelerx_b200_$v.so
  timeout 300 python bench.py --no-cpu-baseline --steps 4 --reactor ReactorRK64 2>>gpurun_out/${TAG}_bench.err > gpurun_out/${TAG}_bench_lidryer_rk64_128_$v.json
  timeout 300 python bench.py --no-cpu-baseline --steps 4 --reactor ReactorRK64 --states pmf 2>>gpurun_out/${TAG}_bench.err > gpurun_out/${TAG}_bench_lidryer_rk64_128_pmf_$v.json
done
unset PRX_LIB_PATH
timeout 300 python bench.py --no-cpu-baseline --steps 4 --reactor ReactorRK64 --states pmf 2>>gpurun_out/${TAG}_bench.err > gpurun_out/${TAG}_bench_lidryer_rk64_128_pmf_default.json
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/${TAG}_bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], "%.3e %s" % (d["value"], d["unit"]), d.get("ms_per_step"), (d.get("roofline") or {}).get("frac"), d.get("failed_cells"))
    except Exception as e:
        print(f, "unreadable", e)
PY
