#!/bin/bash
# GPU session 7: gen-1 kernel with staged LU: barrier modes, out-of-line transcendentals (A/B libs), variants parity
set -u
mkdir -p gpurun_out
echo "== pytest -m gpu"
timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 | tee gpurun_out/pytest_gpu7.log
OUT=gpurun_out/sweep7.jsonl
: > $OUT
run() { local tag=$1; shift
  echo "== $tag"
  ( env "$@" timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline $BARGS 2>>gpurun_out/sweep7.err | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); d['tag'] = '$tag'; print(json.dumps(d))
" >> $OUT )
}
NIM=$PWD/pelec_b200/lib/libpelerx_b200_nim.so
BARGS=""
run lidryer_v0_b0 PRX_BDF_VARIANT=0 PRX_BLOCK_WARPS=0
run lidryer_v0_b1 PRX_BDF_VARIANT=0 PRX_BLOCK_WARPS=1
run lidryer_v0_b3 PRX_BDF_VARIANT=0 PRX_BLOCK_WARPS=3
run lidryer_v0_b0_nim PRX_BDF_VARIANT=0 PRX_BLOCK_WARPS=0 PRX_LIB_PATH=$NIM
run lidryer_v0_b1_nim PRX_BDF_VARIANT=0 PRX_BLOCK_WARPS=1 PRX_LIB_PATH=$NIM
run lidryer_v0_b3_nim PRX_BDF_VARIANT=0 PRX_BLOCK_WARPS=3 PRX_LIB_PATH=$NIM
run lidryer_v3_nim PRX_BDF_VARIANT=3 PRX_LIB_PATH=$NIM
BARGS="--reactor ReactorRK64"
run lidryer_rk64 PRX_BLOCK_WARPS=0
run lidryer_rk64_nim PRX_LIB_PATH=$NIM
BARGS="--mech drm19 --n 64"
run drm19_v0 PRX_BDF_VARIANT=0
run drm19_v0_nim PRX_BDF_VARIANT=0 PRX_LIB_PATH=$NIM
run drm19_v0_b1_nim PRX_BDF_VARIANT=0 PRX_BLOCK_WARPS=1 PRX_LIB_PATH=$NIM
python - <<PY
import json
for l in open("$OUT"):
    d=json.loads(l); print(d["tag"], "%.3e cells/s"%d["value"], "%.1f ms"%d["ms_per_step"], "frac %.4f"%d["roofline"]["frac"], "e2e %.3e"%d["e2e"]["value"])
PY
tail -5 gpurun_out/sweep7.err
echo "== ncu full (bdf v0, LiDryer 64^3, std lib)"
PRX_BDF_VARIANT=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:bdf_kernel -s 1 -c 1 -f -o gpurun_out/prof_bdf0b_lidryer \
   python bench.py --steps 1 --warmup 3 --n 64 --no-cpu-baseline > gpurun_out/ncu_full7.log 2>&1
tail -2 gpurun_out/ncu_full7.log | cut -c1-200
echo "== ncu full (bdf v0, LiDryer 64^3, nim lib)"
PRX_LIB_PATH=$NIM PRX_BDF_VARIANT=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:bdf_kernel -s 1 -c 1 -f -o gpurun_out/prof_bdf0c_lidryer \
   python bench.py --steps 1 --warmup 3 --n 64 --no-cpu-baseline > gpurun_out/ncu_full7b.log 2>&1
tail -2 gpurun_out/ncu_full7b.log | cut -c1-200
