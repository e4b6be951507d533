#!/bin/bash
# session r02z: variant 6 (integrator vectors in an L2-pinned global buffer, gen-1 block shape): parity + timing
mkdir -p gpurun_out
L=pelec_b200/lib/libpelerx_b200_v6.so
PRX_LIB_PATH=$L python -m pytest tests/test_gpu_parity.py -m gpu -q -k "generation and LiDryer and 6" 2>&1 | tail -5
for cfg in "0 1" "6 1" "6 0"; do set -- $cfg
  PRX_LIB_PATH=$L PRX_BDF_VARIANT=$1 PRX_L2_PERSIST=$2 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-ref-gpu --strong none > gpurun_out/r02z_bench_lidryer_v$1_p$2.json 2> gpurun_out/r02z.err
  python - <<P
import json
try:
    d=json.loads(open("gpurun_out/r02z_bench_lidryer_v$1_p$2.json").read().strip().splitlines()[-1])
    print("variant $1 persist $2: %.2f ms, %s, failed %s"%(d["ms_per_step"], d["roofline"]["kernel"], d["failed_cells"]))
except Exception as e:
    print("variant $1 FAILED", e); print(open("gpurun_out/r02z.err").read()[-400:])
P
done
