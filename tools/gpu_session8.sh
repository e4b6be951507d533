#!/bin/bash
# GPU session 8: block-size / occupancy A/B builds of the gen-1 kernel, Newton matrix in global memory for dodecane
set -u
mkdir -p gpurun_out
echo "== pytest -m gpu"
timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/pytest_gpu8.log
OUT=gpurun_out/sweep8.jsonl
: > $OUT
run() { local tag=$1; shift
  echo "== $tag"
  ( env "$@" timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline $BARGS 2>>gpurun_out/sweep8.err | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); d['tag'] = '$tag'; print(json.dumps(d))
" >> $OUT )
}
L=$PWD/pelec_b200/lib/libpelerx_b200
BARGS="--mech dodecane_lu --n 32"
run dodecane_g32_aglobal PRX_BDF_VARIANT=32
run dodecane_g32_asmem PRX_BDF_VARIANT=32 PRX_COOP_A_GLOBAL=0
BARGS=""
for t in b256m2 b128m4 b128m6 b64m8 b256m1; do
run lidryer_${t}_b0 PRX_BDF_VARIANT=0 PRX_LIB_PATH=${L}_$t.so
run lidryer_${t}_b1 PRX_BDF_VARIANT=0 PRX_BLOCK_WARPS=1 PRX_LIB_PATH=${L}_$t.so
done
BARGS="--mech drm19 --n 64"
for t in b256m2 d128m2 d128m4; do
run drm19_${t}_b0 PRX_BDF_VARIANT=0 PRX_LIB_PATH=${L}_$t.so
run drm19_${t}_b1 PRX_BDF_VARIANT=0 PRX_BLOCK_WARPS=1 PRX_LIB_PATH=${L}_$t.so
done
python - <<PY
import json
for l in open("$OUT"):
    d=json.loads(l); print(d["tag"], "%.3e cells/s"%d["value"], "%.1f ms"%d["ms_per_step"], "frac %.4f"%d["roofline"]["frac"], "e2e %.3e"%d["e2e"]["value"])
PY
tail -5 gpurun_out/sweep8.err
