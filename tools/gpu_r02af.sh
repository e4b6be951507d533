#!/bin/bash
# session r02af: (a) external species source in a shared-memory column (PRX_SMEM_SRC A/B library), (b) hottest cells first (PRX_SORT_CELLS=2)
mkdir -p gpurun_out
PRX_LIB_PATH=pelec_b200/lib/libpelerx_b200_smemsrc.so python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "LiDryer and (golden or full_size)" 2>&1 | tail -4 | tee gpurun_out/r02af_pytest_smemsrc.log
run() { # tag lib sort
  PRX_SORT_CELLS=$3 PRX_LIB_PATH=$2 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-ref-gpu --strong none > gpurun_out/r02af_bench_lidryer_$1.json 2> gpurun_out/r02af_$1.err
  python - <<P
import json
try:
    d=json.loads(open("gpurun_out/r02af_bench_lidryer_$1.json").read().strip().splitlines()[-1])
    print("$1: %.2f ms, e2e %.2f ms, failed %s"%(d["ms_per_step"], d["e2e"]["ms_per_step"], d["failed_cells"]))
except Exception as e:
    print("$1 FAILED", e); print(open("gpurun_out/r02af_$1.err").read()[-300:])
P
}
B=pelec_b200/lib/libpelerx_b200.so
S=pelec_b200/lib/libpelerx_b200_smemsrc.so
run base_1 $B 1
run smemsrc_1 $S 1
run desc_1 $B 2
run base_2 $B 1
run smemsrc_2 $S 1
run desc_2 $B 2
run smemsrc_desc $S 2
