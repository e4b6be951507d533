#!/bin/bash
# quick GPU iteration: parity of the default kernels + bench lines of the three mechanisms (no CPU baseline)
set -u
TAG=${1:-x}
mkdir -p gpurun_out
for m in LiDryer drm19; do timeout 300 python tools/dbg.py $m 2>&1 | tail -1; done
OUT=gpurun_out/quick_$TAG.jsonl
: > $OUT
timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline >> $OUT 2>gpurun_out/quick_$TAG.err
timeout 300 python bench.py --steps 3 --warmup 3 --mech drm19 --n 64 --no-cpu-baseline >> $OUT 2>>gpurun_out/quick_$TAG.err
if [ "${2:-}" = "all" ]; then
timeout 300 python bench.py --steps 3 --warmup 3 --reactor ReactorRK64 --no-cpu-baseline >> $OUT 2>>gpurun_out/quick_$TAG.err
timeout 600 python bench.py --steps 3 --warmup 3 --mech dodecane_lu --n 32 --no-cpu-baseline >> $OUT 2>>gpurun_out/quick_$TAG.err
fi
python - <<PY
import json
for l in open("$OUT"):
    d=json.loads(l); print(d["config"]["mechanism"], d["config"]["reactor"], "%.3e cells/s"%d["value"], "%.1f ms"%d["ms_per_step"], "frac %.4f"%d["roofline"]["frac"], "e2e %.3e"%d["e2e"]["value"])
PY
tail -5 gpurun_out/quick_$TAG.err
