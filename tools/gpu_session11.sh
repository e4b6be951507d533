#!/bin/bash
# GPU session 11: column-batched cooperative LU / solves: parity + dodecane
set -u
mkdir -p gpurun_out
echo "== pytest -m gpu"
timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/pytest_gpu11.log
OUT=gpurun_out/sweep11.jsonl
: > $OUT
run() { local tag=$1; shift
  echo "== $tag"
  ( env "$@" timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline $BARGS 2>>gpurun_out/sweep11.err | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); d['tag'] = '$tag'; print(json.dumps(d))
" >> $OUT )
}
BARGS="--mech dodecane_lu --n 32"
run dodecane_g32 PRX_BDF_VARIANT=32
run dodecane_g32_asmem PRX_BDF_VARIANT=32 PRX_COOP_A_GLOBAL=0
run dodecane_g32_w16 PRX_BDF_VARIANT=32 PRX_MAX_WARPS_PER_SM=16
BARGS=""
run lidryer_v0 PRX_BDF_VARIANT=0
run lidryer_g8 PRX_BDF_VARIANT=8
BARGS="--mech drm19 --n 64"
run drm19_v0 PRX_BDF_VARIANT=0
run drm19_g16 PRX_BDF_VARIANT=16
run drm19_g8 PRX_BDF_VARIANT=8
python - <<PY
import json
for l in open("$OUT"):
    d=json.loads(l); print(d["tag"], "%.3e cells/s"%d["value"], "%.1f ms"%d["ms_per_step"], "frac %.4f"%d["roofline"]["frac"], "e2e %.3e"%d["e2e"]["value"])
PY
tail -5 gpurun_out/sweep11.err
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:bdfc_kernel -s 1 -c 1 -f -o gpurun_out/prof_bdfc2_dodecane \
   python bench.py --steps 1 --warmup 2 --mech dodecane_lu --n 16 --no-cpu-baseline > gpurun_out/ncu_full11.log 2>&1
tail -2 gpurun_out/ncu_full11.log | cut -c1-200
