#!/bin/bash
set -u
mkdir -p gpurun_out
OUT=gpurun_out/sweep18.jsonl
: > $OUT
run() { local tag=$1; shift
  ( env "$@" timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline $BARGS 2>>gpurun_out/sweep18.err | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); d['tag'] = '$tag'; print(json.dumps(d))
" >> $OUT )
}
for b in 0 1 2 3; do
BARGS=""; run uni_l_b$b PRX_BLOCK_WARPS=$b
BARGS="--states pmf"; run pmf_l_b$b PRX_BLOCK_WARPS=$b
BARGS="--mech drm19 --n 64"; run uni_d_b$b PRX_BLOCK_WARPS=$b
done
BARGS=""; run uni_l_v3 PRX_BDF_VARIANT=3
BARGS=""; run uni_l_v3_W4 PRX_BDF_VARIANT=3 PRX_BLOCK_WARPS=4
BARGS=""; run uni_l_v2 PRX_BDF_VARIANT=2
BARGS=""; run uni_l_g8 PRX_BDF_VARIANT=8
python - <<PY
import json
for l in open("$OUT"):
    d=json.loads(l); print(d["tag"], "%.3e cells/s"%d["value"], "%.1f ms"%d["ms_per_step"], "e2e %.3e"%d["e2e"]["value"])
PY
tail -3 gpurun_out/sweep18.err
