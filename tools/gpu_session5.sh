#!/bin/bash
# GPU session 5: hot/cold split thread-per-cell kernel (bdf3): parity + sweep + ncu
set -u
mkdir -p gpurun_out
for v in 3 4; do echo "== dbg LiDryer variant $v"; PRX_BDF_VARIANT=$v timeout 600 python tools/dbg.py LiDryer 2>&1 | tail -2; done
echo "== dbg drm19 variant 3"; PRX_BDF_VARIANT=3 timeout 600 python tools/dbg.py drm19 2>&1 | tail -2
OUT=gpurun_out/sweep5.jsonl
: > $OUT
run() { local tag=$1; shift
  echo "== $tag"
  ( env "$@" timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline $BARGS 2>>gpurun_out/sweep5.err | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); d['tag'] = '$tag'; print(json.dumps(d))
" >> $OUT )
}
BARGS=""
run lidryer_v3 PRX_BDF_VARIANT=3
run lidryer_v4 PRX_BDF_VARIANT=4
run lidryer_v3_w8 PRX_BDF_VARIANT=3 PRX_MAX_WARPS_PER_SM=8
run lidryer_v3_w4 PRX_BDF_VARIANT=3 PRX_MAX_WARPS_PER_SM=4
run lidryer_v0 PRX_BDF_VARIANT=0
BARGS="--mech drm19 --n 64"
run drm19_v3 PRX_BDF_VARIANT=3
run drm19_v4 PRX_BDF_VARIANT=4
python - <<PY
import json
for l in open("$OUT"):
    d=json.loads(l); print(d["tag"], "%.3e cells/s"%d["value"], "%.1f ms"%d["ms_per_step"], "frac %.4f"%d["roofline"]["frac"], "e2e %.3e"%d["e2e"]["value"])
PY
tail -5 gpurun_out/sweep5.err
echo "== ncu full (bdf3, LiDryer 64^3)"
PRX_BDF_VARIANT=3 timeout 900 ncu --set full --clock-control none --import-source on -k regex:bdf3_kernel -s 1 -c 1 -f -o gpurun_out/prof_bdf3_lidryer \
   python bench.py --steps 1 --warmup 3 --n 64 --no-cpu-baseline > gpurun_out/ncu_full5.log 2>&1
tail -2 gpurun_out/ncu_full5.log | cut -c1-200
echo "== ncu full (bdf v0 with block barrier, LiDryer 64^3)"
PRX_BDF_VARIANT=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:bdf_kernel -s 1 -c 1 -f -o gpurun_out/prof_bdf0_lidryer \
   python bench.py --steps 1 --warmup 3 --n 64 --no-cpu-baseline > gpurun_out/ncu_full5b.log 2>&1
tail -2 gpurun_out/ncu_full5b.log | cut -c1-200
