#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest -m gpu"
timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/pytest_gpu13.log
OUT=gpurun_out/sweep13.jsonl
: > $OUT
run() { local tag=$1; shift
  echo "== $tag"
  ( env "$@" timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline $BARGS 2>>gpurun_out/sweep13.err | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); d['tag'] = '$tag'; print(json.dumps(d))
" >> $OUT )
}
BARGS=""
run lidryer_b0 PRX_BLOCK_WARPS=0
run lidryer_b2 PRX_BLOCK_WARPS=2
BARGS="--mech drm19 --n 64"
run drm19_b0 PRX_BLOCK_WARPS=0
run drm19_b2 PRX_BLOCK_WARPS=2
BARGS="--mech dodecane_lu --n 32"
run dodecane PRX_BLOCK_WARPS=0
python - <<PY
import json
for l in open("$OUT"):
    d=json.loads(l); print(d["tag"], "%.3e cells/s"%d["value"], "%.1f ms"%d["ms_per_step"], "frac %.4f"%d["roofline"]["frac"], "e2e %.3e"%d["e2e"]["value"])
PY
tail -5 gpurun_out/sweep13.err
