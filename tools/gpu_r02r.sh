#!/bin/bash
# session r02r: storage variants re-measured on the current code (full-variants A/B library), PeleC application drop-in test
mkdir -p gpurun_out
python -m pytest tests/test_pelec_dropin.py -m gpu -x -q -s 2>&1 | tail -8 > gpurun_out/r02r_pytest_pelec.log
cat gpurun_out/r02r_pytest_pelec.log
for v in 0 1 2 3 4 5; do
  PRX_LIB_PATH=pelec_b200/lib/libpelerx_b200_full.so PRX_BDF_VARIANT=$v timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-ref-gpu --strong none > gpurun_out/r02r_bench_lidryer_v$v.json 2> gpurun_out/r02r_v$v.err
  python - <<P
import json
try:
    d=json.loads(open("gpurun_out/r02r_bench_lidryer_v$v.json").read().strip().splitlines()[-1])
    print("variant $v", d["roofline"]["kernel"], "%.1f ms"%d["ms_per_step"], "failed", d["failed_cells"])
except Exception as e:
    print("variant $v FAILED", e); print(open("gpurun_out/r02r_v$v.err").read()[-300:])
P
done
