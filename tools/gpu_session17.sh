#!/bin/bash
# GPU session 17: temperature-sorted cell order: parity + divergent and uniform workloads
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
OUT=gpurun_out/sweep17.jsonl
: > $OUT
run() { local tag=$1; shift
  ( env "$@" timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline $BARGS 2>>gpurun_out/sweep17.err | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); d['tag'] = '$tag'; print(json.dumps(d))
" >> $OUT )
}
BARGS="--states pmf"
run pmf_l_sort
run pmf_l_nosort PRX_SORT_CELLS=0
run pmf_l_sort_r1 PRX_REFILL_MIN=1
BARGS="--states pmf --rtol 1e-6 --max-order 2"
run pmf_l_pelec_sort
BARGS="--states pmf --reactor ReactorRK64"
run pmf_l_rk64_sort
run pmf_l_rk64_nosort PRX_SORT_CELLS=0
BARGS="--states pmf --mech drm19 --n 64"
run pmf_d_sort
run pmf_d_nosort PRX_SORT_CELLS=0
BARGS=""
run uni_l_sort
run uni_l_nosort PRX_SORT_CELLS=0
BARGS="--mech drm19 --n 64"
run uni_d_sort
BARGS="--mech dodecane_lu --n 32"
run uni_dod_sort
BARGS="--reactor ReactorRK64"
run uni_l_rk64_sort
python - <<PY
import json
for l in open("$OUT"):
    d=json.loads(l); print(d["tag"], "%.3e cells/s"%d["value"], "%.1f ms"%d["ms_per_step"], "frac %.4f"%d["roofline"]["frac"], "e2e %.3e"%d["e2e"]["value"])
PY
tail -3 gpurun_out/sweep17.err
