#!/bin/bash
set -u
mkdir -p gpurun_out
OUT=gpurun_out/sweep20.jsonl
: > $OUT
run() { local tag=$1; shift
  ( env "$@" timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline $BARGS 2>>gpurun_out/sweep20.err | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); d['tag'] = '$tag'; print(json.dumps(d))
" >> $OUT )
}
L=$PWD/pelec_b200/lib/libpelerx_b200
for t in "$@"; do
PRX_LIB_PATH=${L}_$t.so timeout 300 python tools/dbg.py LiDryer 2>&1 | tail -1
BARGS=""; run uni_l_$t PRX_LIB_PATH=${L}_$t.so
BARGS="--states pmf"; run pmf_l_$t PRX_LIB_PATH=${L}_$t.so
BARGS="--mech drm19 --n 64"; run uni_d_$t PRX_LIB_PATH=${L}_$t.so
done
python - <<PY
import json
for l in open("$OUT"):
    d=json.loads(l); print(d["tag"], "%.3e cells/s"%d["value"], "%.1f ms"%d["ms_per_step"], "e2e %.3e"%d["e2e"]["value"])
PY
tail -3 gpurun_out/sweep20.err
