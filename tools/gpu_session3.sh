#!/bin/bash
# GPU session 3: cooperative (G lanes per cell) BDF kernel: parity + G sweep + ncu
set -u
mkdir -p gpurun_out
TAG=${1:-3}
echo "== quick parity (dbg.py) LiDryer / drm19"
timeout 300 python tools/dbg.py LiDryer 2>&1 | tail -3
timeout 300 python tools/dbg.py drm19 2>&1 | tail -3
echo "== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu$TAG.log
OUT=gpurun_out/sweep$TAG.jsonl
: > $OUT
run() { # tag, env..., -- bench args
  local tag=$1; shift
  echo "== $tag"
  ( env "$@" timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline $BARGS 2>>gpurun_out/sweep$TAG.err | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); d['tag'] = '$tag'; print(json.dumps(d))
" >> $OUT )
}
BARGS=""
run lidryer_g8 PRX_BDF_VARIANT=8
run lidryer_g16 PRX_BDF_VARIANT=16
run lidryer_g32 PRX_BDF_VARIANT=32
run lidryer_g8_w16 PRX_BDF_VARIANT=8 PRX_MAX_WARPS_PER_SM=16
BARGS="--mech drm19 --n 64"
run drm19_g16 PRX_BDF_VARIANT=16
run drm19_g32 PRX_BDF_VARIANT=32
run drm19_g8 PRX_BDF_VARIANT=8
python - <<PY
import json
for l in open("$OUT"):
    d=json.loads(l); print(d["tag"], "%.3e cells/s"%d["value"], "%.1f ms"%d["ms_per_step"], "frac %.4f"%d["roofline"]["frac"], "e2e %.3e"%d["e2e"]["value"], d["counters_per_cell"])
PY
tail -5 gpurun_out/sweep$TAG.err
echo "== ncu full (bdfc, LiDryer 64^3)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:bdfc_kernel -s 1 -c 1 -f -o gpurun_out/prof_bdfc_lidryer \
   python bench.py --steps 1 --warmup 3 --n 64 --no-cpu-baseline > gpurun_out/ncu_full$TAG.log 2>&1
tail -3 gpurun_out/ncu_full$TAG.log | cut -c1-300
