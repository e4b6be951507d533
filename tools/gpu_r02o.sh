#!/bin/bash
# Round 2, session O: consolidated build - full GPU suite, dodecane re-check, EOS kernel with 128-bit loads
set -u
TAG=r02o
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee gpurun_out/${TAG}_pytest_gpu.log
for i in 1 2; do
timeout 600 python bench.py --no-cpu-baseline --mech dodecane_lu --n 32 --steps 2 --warmup 1 2>>gpurun_out/${TAG}_bench.err > gpurun_out/${TAG}_bench_dodecane_32_run$i.json
done
timeout 300 python tools/bench_eos.py > gpurun_out/${TAG}_eos_lidryer_256.json 2>gpurun_out/${TAG}_eos.err; cat gpurun_out/${TAG}_eos_lidryer_256.json | cut -c1-300
timeout 300 python tools/bench_eos.py --mech drm19 > gpurun_out/${TAG}_eos_drm19_256.json 2>>gpurun_out/${TAG}_eos.err; cat gpurun_out/${TAG}_eos_drm19_256.json | cut -c1-300
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/${TAG}_bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], "%.3e %s" % (d["value"], d["unit"]), d.get("ms_per_step"), (d.get("roofline") or {}).get("frac"), d.get("failed_cells"))
    except Exception as e:
        print(f, "unreadable", e)
PY
