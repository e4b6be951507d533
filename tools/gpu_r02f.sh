#!/bin/bash
# Round 2, session F: A/B - Pascal-triangle predict (p1/x1) vs statement form (p0/x0); paired exponentials (x*) vs single (p*)
set -u
TAG=r02f
mkdir -p gpurun_out
for v in p0 p1 x0 x1; do
  export PRX_LIB_PATH=$PWD/pelec_b200/lib/libpelerx_b200_$v.so
  timeout 300 python bench.py --no-cpu-baseline --steps 4 --strong none 2>>gpurun_out/${TAG}_bench.err > gpurun_out/${TAG}_bench_lidryer_cvode_128_$v.json
  timeout 300 python bench.py --no-cpu-baseline --steps 4 --reactor ReactorRK64 2>>gpurun_out/${TAG}_bench.err > gpurun_out/${TAG}_bench_lidryer_rk64_128_$v.json
done
export PRX_LIB_PATH=$PWD/pelec_b200/lib/libpelerx_b200_x1.so
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "LiDryer or lidryer or subbox or flatten" 2>&1 | tail -5
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
