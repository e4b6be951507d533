#!/bin/bash
# GPU session 16: divergent workload (flame-table states): barrier modes, batched refill, lane-level deferral
set -u
mkdir -p gpurun_out
OUT=gpurun_out/sweep16.jsonl
: > $OUT
run() { local tag=$1; shift
  ( env "$@" timeout 600 python bench.py --states pmf --steps 3 --warmup 2 --no-cpu-baseline $BARGS 2>>gpurun_out/sweep16.err | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); d['tag'] = '$tag'; print(json.dumps(d))
" >> $OUT )
}
L=$PWD/pelec_b200/lib/libpelerx_b200
BARGS=""
run l_is1_b0_r1
run l_is1_b1_r1 PRX_BLOCK_WARPS=1
run l_is1_b0_r8 PRX_REFILL_MIN=8
run l_is1_b0_r16 PRX_REFILL_MIN=16
run l_is1_b0_r32 PRX_REFILL_MIN=32
run l_is1_b1_r16 PRX_BLOCK_WARPS=1 PRX_REFILL_MIN=16
run l_is1_b1_r32 PRX_BLOCK_WARPS=1 PRX_REFILL_MIN=32
run l_is0_b0_r1 PRX_LIB_PATH=${L}_is0.so
run l_is0_b0_d8 PRX_LIB_PATH=${L}_is0.so PRX_JAC_DEFER=8 PRX_LU_DEFER=8
run l_is0_b0_d16_r16 PRX_LIB_PATH=${L}_is0.so PRX_JAC_DEFER=16 PRX_LU_DEFER=16 PRX_REFILL_MIN=16
run l_is0_b1_d8 PRX_LIB_PATH=${L}_is0.so PRX_JAC_DEFER=8 PRX_LU_DEFER=8 PRX_BLOCK_WARPS=1
run l_g8 PRX_BDF_VARIANT=8
run l_v3 PRX_BDF_VARIANT=3
BARGS="--mech drm19 --n 64"
run d_is1_b0_r1
run d_is1_b1_r1 PRX_BLOCK_WARPS=1
run d_is1_b0_r16 PRX_REFILL_MIN=16
run d_is1_b0_r32 PRX_REFILL_MIN=32
python - <<PY
import json
for l in open("$OUT"):
    d=json.loads(l); print(d["tag"], "%.3e cells/s"%d["value"], "%.1f ms"%d["ms_per_step"], "e2e %.3e"%d["e2e"]["value"])
PY
tail -3 gpurun_out/sweep16.err
echo "== uniform workload sanity with refill 32"
PRX_REFILL_MIN=32 timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('reacteval r32', d['ms_per_step'])"
