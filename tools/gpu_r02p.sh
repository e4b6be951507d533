#!/bin/bash
# session r02p: the reference's own GPU path (oracle/ref_gpu) on the B200
mkdir -p gpurun_out
run() { # name args...
  name=$1; shift
  timeout 900 python tools/ref_gpu_bench.py "$@" > gpurun_out/r02p_refgpu_$name.json 2> gpurun_out/r02p_refgpu_$name.err
  python - <<P
import json
try:
    d=json.loads(open("gpurun_out/r02p_refgpu_$name.json").read().strip().splitlines()[-1])
    print("$name", d.get("ms_per_call"), "%.3g cells/s"%d["value"], d.get("react_return"), d.get("check"))
except Exception as e:
    print("$name FAILED", e); print(open("gpurun_out/r02p_refgpu_$name.err").read()[-400:])
P
}
for n in 16 32 64 128; do run lidryer_gmres_$n --mech LiDryer --n $n --solve-type GMRES --check --reps 2; done
run lidryer_gmres_128_pelec_tols --mech LiDryer --n 128 --solve-type GMRES --rtol 1e-6 --atol 1e-10 --max-order 2 --check --reps 2
for n in 64 128; do run lidryer_rk64_$n --mech LiDryer --n $n --reactor ReactorRK64 --check --reps 2; done
