#!/bin/bash
# GPU session 6: register-staged LU/solve in the thread-per-cell kernels, phase-aligned blocks
set -u
mkdir -p gpurun_out
for v in 3 0 2; do echo "== dbg LiDryer variant $v"; PRX_BDF_VARIANT=$v timeout 600 python tools/dbg.py LiDryer 2>&1 | tail -1; done
echo "== dbg LiDryer variant 3 W=4"; PRX_BDF_VARIANT=3 PRX_BLOCK_WARPS=4 timeout 600 python tools/dbg.py LiDryer 2>&1 | tail -1
echo "== dbg drm19 variant 3"; PRX_BDF_VARIANT=3 timeout 600 python tools/dbg.py drm19 2>&1 | tail -1
OUT=gpurun_out/sweep6.jsonl
: > $OUT
run() { local tag=$1; shift
  echo "== $tag"
  ( env "$@" timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline $BARGS 2>>gpurun_out/sweep6.err | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); d['tag'] = '$tag'; print(json.dumps(d))
" >> $OUT )
}
BARGS=""
run lidryer_v3 PRX_BDF_VARIANT=3
run lidryer_v3_W2 PRX_BDF_VARIANT=3 PRX_BLOCK_WARPS=2
run lidryer_v3_W4 PRX_BDF_VARIANT=3 PRX_BLOCK_WARPS=4
run lidryer_v4 PRX_BDF_VARIANT=4
run lidryer_v4_W4 PRX_BDF_VARIANT=4 PRX_BLOCK_WARPS=4
run lidryer_v2 PRX_BDF_VARIANT=2
run lidryer_v1 PRX_BDF_VARIANT=1
run lidryer_v0 PRX_BDF_VARIANT=0
BARGS="--mech drm19 --n 64"
run drm19_v3 PRX_BDF_VARIANT=3
run drm19_v0 PRX_BDF_VARIANT=0
python - <<PY
import json
for l in open("$OUT"):
    d=json.loads(l); print(d["tag"], "%.3e cells/s"%d["value"], "%.1f ms"%d["ms_per_step"], "frac %.4f"%d["roofline"]["frac"], "e2e %.3e"%d["e2e"]["value"])
PY
tail -5 gpurun_out/sweep6.err
echo "== ncu full (bdf3, LiDryer 64^3)"
PRX_BDF_VARIANT=3 timeout 900 ncu --set full --clock-control none --import-source on -k regex:bdf3_kernel -s 1 -c 1 -f -o gpurun_out/prof_bdf3b_lidryer \
   python bench.py --steps 1 --warmup 3 --n 64 --no-cpu-baseline > gpurun_out/ncu_full6.log 2>&1
tail -2 gpurun_out/ncu_full6.log | cut -c1-200
