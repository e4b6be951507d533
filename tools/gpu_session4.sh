#!/bin/bash
# GPU session 4: cooperative kernel with the cooperative table-driven Jacobian, all three mechanisms
set -u
mkdir -p gpurun_out
for m in LiDryer drm19 dodecane_lu; do
echo "== dbg $m"; timeout 600 python tools/dbg.py $m 2>&1 | tail -2
done
echo "== pytest -m gpu"
timeout 2000 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu4.log
OUT=gpurun_out/sweep4.jsonl
: > $OUT
run() { local tag=$1; shift
  echo "== $tag"
  ( env "$@" timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline $BARGS 2>>gpurun_out/sweep4.err | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); d['tag'] = '$tag'; print(json.dumps(d))
" >> $OUT )
}
BARGS="--mech dodecane_lu --n 32"
run dodecane_g32 PRX_BDF_VARIANT=32
run dodecane_g16 PRX_BDF_VARIANT=16
BARGS=""
run lidryer_g8 PRX_BDF_VARIANT=8
BARGS="--mech drm19 --n 64"
run drm19_g8 PRX_BDF_VARIANT=8
python - <<PY
import json
for l in open("$OUT"):
    d=json.loads(l); print(d["tag"], "%.3e cells/s"%d["value"], "%.1f ms"%d["ms_per_step"], "frac %.4f"%d["roofline"]["frac"], "e2e %.3e"%d["e2e"]["value"], d["counters_per_cell"])
PY
tail -5 gpurun_out/sweep4.err
