#!/bin/bash
# session r02ag: gen-1 BDF with hot state in shared-memory columns: level 1 = external source, 2 = + ewt, 3 = + acor (A/B libraries)
mkdir -p gpurun_out
for L in 2 3; do
PRX_LIB_PATH=pelec_b200/lib/libpelerx_b200_smem$L.so python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "LiDryer and (golden or full_size)" 2>&1 | tail -2 | tee gpurun_out/r02ag_pytest_smem$L.log
done
run() { # tag lib
  PRX_LIB_PATH=$2 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-ref-gpu --strong none > gpurun_out/r02ag_bench_lidryer_$1.json 2> gpurun_out/r02ag_$1.err
  python - <<P
import json
try:
    d=json.loads(open("gpurun_out/r02ag_bench_lidryer_$1.json").read().strip().splitlines()[-1])
    print("$1: %.2f ms, e2e %.2f ms, failed %s"%(d["ms_per_step"], d["e2e"]["ms_per_step"], d["failed_cells"]))
except Exception as e:
    print("$1 FAILED", e); print(open("gpurun_out/r02ag_$1.err").read()[-300:])
P
}
for rep in 1 2; do
run base_$rep pelec_b200/lib/libpelerx_b200.so
for L in 1 2 3; do run smem${L}_$rep pelec_b200/lib/libpelerx_b200_smem$L.so; done
done
