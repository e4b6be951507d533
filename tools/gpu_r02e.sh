#!/bin/bash
# Round 2, session E: full GPU test suite (new: flat / ignition / dodecane / react_state / multi-box) + default bench + small boxes
set -u
TAG=r02e
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 | tee gpurun_out/${TAG}_pytest_gpu.log
timeout 300 python bench.py --no-cpu-baseline --strong none 2>>gpurun_out/${TAG}_bench.err > gpurun_out/${TAG}_bench_lidryer_cvode_128.json
timeout 300 python bench.py --no-cpu-baseline --reactor ReactorRK64 2>>gpurun_out/${TAG}_bench.err > gpurun_out/${TAG}_bench_lidryer_rk64_128.json
timeout 300 python bench.py --no-cpu-baseline --mech drm19 --n 64 2>>gpurun_out/${TAG}_bench.err > gpurun_out/${TAG}_bench_drm19_cvode_64.json
timeout 600 python bench.py --no-cpu-baseline --strong drm19:128:32 --strong-steps 1 2>>gpurun_out/${TAG}_bench.err > gpurun_out/${TAG}_bench_lidryer_cvode_128_strong.json
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/${TAG}_bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], "%.3e %s" % (d["value"], d["unit"]), d.get("ms_per_step"), (d.get("roofline") or {}).get("frac"), (d.get("e2e") or {}).get("value"), d.get("failed_cells"), d.get("strong"))
    except Exception as e:
        print(f, "unreadable", e)
PY
tail -5 gpurun_out/${TAG}_bench.err
