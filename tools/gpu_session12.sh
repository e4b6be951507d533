#!/bin/bash
set -u
mkdir -p gpurun_out
OUT=gpurun_out/sweep12.jsonl
: > $OUT
run() { local tag=$1; shift
  echo "== $tag"
  ( env "$@" timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline $BARGS 2>>gpurun_out/sweep12.err | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); d['tag'] = '$tag'; print(json.dumps(d))
" >> $OUT )
}
L=$PWD/pelec_b200/lib/libpelerx_b200
for t in "$@"; do
echo "== dbg $t"; PRX_LIB_PATH=${L}_$t.so timeout 600 python tools/dbg.py LiDryer 2>&1 | tail -1
PRX_LIB_PATH=${L}_$t.so timeout 600 python tools/dbg.py drm19 2>&1 | tail -1
BARGS=""
run lidryer_$t PRX_LIB_PATH=${L}_$t.so
BARGS="--mech drm19 --n 64"
run drm19_$t PRX_LIB_PATH=${L}_$t.so
done
python - <<PY
import json
for l in open("$OUT"):
    d=json.loads(l); print(d["tag"], "%.3e cells/s"%d["value"], "%.1f ms"%d["ms_per_step"], "frac %.4f"%d["roofline"]["frac"], "e2e %.3e"%d["e2e"]["value"])
PY
tail -5 gpurun_out/sweep12.err
