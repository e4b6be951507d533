#!/bin/bash
# GPU session 10: final-defaults build: parity, block-level batching of slow requests, profiles
set -u
mkdir -p gpurun_out
echo "== pytest -m gpu"
timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/pytest_gpu10.log
OUT=gpurun_out/sweep10.jsonl
: > $OUT
run() { local tag=$1; shift
  echo "== $tag"
  ( env "$@" timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline $BARGS 2>>gpurun_out/sweep10.err | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); d['tag'] = '$tag'; print(json.dumps(d))
" >> $OUT )
}
BARGS=""
run lidryer_d0 PRX_LU_DEFER_BLOCK=0
run lidryer_d25 PRX_LU_DEFER_BLOCK=25
run lidryer_d50 PRX_LU_DEFER_BLOCK=50
run lidryer_d75 PRX_LU_DEFER_BLOCK=75
run lidryer_d50_a12 PRX_LU_DEFER_BLOCK=50 PRX_LU_DEFER_AGE=12
run lidryer_d75_a3 PRX_LU_DEFER_BLOCK=75 PRX_LU_DEFER_AGE=3
BARGS="--mech drm19 --n 64"
run drm19_d0 PRX_LU_DEFER_BLOCK=0
run drm19_d50 PRX_LU_DEFER_BLOCK=50
run drm19_d75_a12 PRX_LU_DEFER_BLOCK=75 PRX_LU_DEFER_AGE=12
BARGS="--mech dodecane_lu --n 32"
run dodecane_g32 PRX_LU_DEFER_BLOCK=0
python - <<PY
import json
for l in open("$OUT"):
    d=json.loads(l); print(d["tag"], "%.3e cells/s"%d["value"], "%.1f ms"%d["ms_per_step"], "frac %.4f"%d["roofline"]["frac"], "e2e %.3e"%d["e2e"]["value"])
PY
tail -5 gpurun_out/sweep10.err
PRX_LU_DEFER_BLOCK=50 timeout 900 ncu --set full --clock-control none --import-source on -k regex:bdf_kernel -s 1 -c 1 -f -o gpurun_out/prof_bdf0e_lidryer \
   python bench.py --steps 1 --warmup 3 --n 64 --no-cpu-baseline > gpurun_out/ncu_full10.log 2>&1
tail -2 gpurun_out/ncu_full10.log | cut -c1-200
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:bdfc_kernel -s 1 -c 1 -f -o gpurun_out/prof_bdfc_dodecane \
   python bench.py --steps 1 --warmup 2 --mech dodecane_lu --n 16 --no-cpu-baseline > gpurun_out/ncu_full10b.log 2>&1
tail -2 gpurun_out/ncu_full10b.log | cut -c1-200
