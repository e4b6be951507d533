#!/bin/bash
# GPU session 2: second-generation BDF kernel (shared-memory columns + table-driven RHS): parity + variant sweep + ncu
set -u
mkdir -p gpurun_out
echo "== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu2.log
OUT=gpurun_out/sweep2.jsonl
: > $OUT
run() { # tag, env..., -- bench args
  local tag=$1; shift
  echo "== $tag"
  ( env "$@" timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline $BARGS 2>>gpurun_out/sweep2.err | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); d['tag'] = '$tag'; print(json.dumps(d))
" >> $OUT )
}
BARGS=""
run lidryer_v0 PRX_BDF_VARIANT=0
run lidryer_v1 PRX_BDF_VARIANT=1
run lidryer_v2 PRX_BDF_VARIANT=2
run lidryer_v2_w4 PRX_BDF_VARIANT=2 PRX_MAX_WARPS_PER_SM=4
run lidryer_v1_w2 PRX_BDF_VARIANT=1 PRX_MAX_WARPS_PER_SM=2
BARGS="--mech drm19 --n 64"
run drm19_v0 PRX_BDF_VARIANT=0
run drm19_v2 PRX_BDF_VARIANT=2
python - <<PY
import json
for l in open("$OUT"):
    d=json.loads(l); print(d["tag"], "%.3e cells/s"%d["value"], "%.1f ms"%d["ms_per_step"], "frac %.4f"%d["roofline"]["frac"], "e2e %.3e"%d["e2e"]["value"], d["counters_per_cell"])
PY
tail -5 gpurun_out/sweep2.err
echo "== ncu full (bdf2, LiDryer 64^3)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:bdf2_kernel -s 1 -c 1 -f -o gpurun_out/prof_bdf2_lidryer \
   python bench.py --steps 1 --warmup 3 --n 64 --no-cpu-baseline > gpurun_out/ncu_full2.log 2>&1
tail -3 gpurun_out/ncu_full2.log
ls -la gpurun_out
