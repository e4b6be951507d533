#!/bin/bash
# GPU session 9: fully unrolled vector statements of the control code (A/B)
set -u
mkdir -p gpurun_out
OUT=gpurun_out/sweep9.jsonl
: > $OUT
run() { local tag=$1; shift
  echo "== $tag"
  ( env "$@" timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline $BARGS 2>>gpurun_out/sweep9.err | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); d['tag'] = '$tag'; print(json.dumps(d))
" >> $OUT )
}
L=$PWD/pelec_b200/lib/libpelerx_b200
for t in vu1 vu32; do
echo "== dbg $t"; PRX_BDF_VARIANT=0 PRX_LIB_PATH=${L}_$t.so timeout 600 python tools/dbg.py LiDryer 2>&1 | tail -1
BARGS=""
run lidryer_${t}_b0 PRX_BDF_VARIANT=0 PRX_LIB_PATH=${L}_$t.so
run lidryer_${t}_b1 PRX_BDF_VARIANT=0 PRX_BLOCK_WARPS=1 PRX_LIB_PATH=${L}_$t.so
BARGS="--mech drm19 --n 64"
run drm19_${t}_b0 PRX_BDF_VARIANT=0 PRX_LIB_PATH=${L}_$t.so
done
python - <<PY
import json
for l in open("$OUT"):
    d=json.loads(l); print(d["tag"], "%.3e cells/s"%d["value"], "%.1f ms"%d["ms_per_step"], "frac %.4f"%d["roofline"]["frac"], "e2e %.3e"%d["e2e"]["value"])
PY
tail -5 gpurun_out/sweep9.err
PRX_LIB_PATH=${L}_vu32.so PRX_BDF_VARIANT=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:bdf_kernel -s 1 -c 1 -f -o gpurun_out/prof_bdf0d_lidryer \
   python bench.py --steps 1 --warmup 3 --n 64 --no-cpu-baseline > gpurun_out/ncu_full9.log 2>&1
tail -2 gpurun_out/ncu_full9.log | cut -c1-200
