#!/bin/bash
# Round 2, session D: A/B of the chemistry code generation (literal pool y/n x custom exp/log10 y/n x inlined math y/n), LiDryer
set -u
TAG=r02d
mkdir -p gpurun_out
for v in a b c d e f; do
  export PRX_LIB_PATH=$PWD/pelec_b200/lib/libpelerx_b200_$v.so
  timeout 300 python bench.py --no-cpu-baseline --steps 3 2>>gpurun_out/${TAG}_bench.err > gpurun_out/${TAG}_bench_lidryer_cvode_128_$v.json
  timeout 300 python bench.py --no-cpu-baseline --steps 3 --reactor ReactorRK64 2>>gpurun_out/${TAG}_bench.err > gpurun_out/${TAG}_bench_lidryer_rk64_128_$v.json
done
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/${TAG}_bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], "%.3e %s" % (d["value"], d["unit"]), d.get("ms_per_step"), (d.get("roofline") or {}).get("frac"), d.get("failed_cells"))
    except Exception as e:
        print(f, "unreadable", e)
PY
tail -3 gpurun_out/${TAG}_bench.err
