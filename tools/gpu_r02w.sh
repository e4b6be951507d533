#!/bin/bash
# session r02w: extra mechanisms (BurkeDryer, FFCM1_Red) on the GPU: parity + throughput
mkdir -p gpurun_out
python -m pytest tests/test_extra_mechanisms.py -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r02w_pytest_extra.log; cat gpurun_out/r02w_pytest_extra.log
for m in BurkeDryer FFCM1_Red; do for rx in ReactorCvode ReactorRK64; do
  python bench.py --mech $m --n 64 --reactor $rx --steps 3 --warmup 3 --no-cpu-baseline --no-ref-gpu --strong none > gpurun_out/r02w_bench_${m}_${rx}_64.json 2> gpurun_out/r02w_bench.err
  python - <<P
import json
try:
    d=json.loads(open("gpurun_out/r02w_bench_${m}_${rx}_64.json").read().strip().splitlines()[-1])
    print("$m $rx 64^3: %.2f ms, %.3g cells/s, roofline %.3f, failed %s"%(d["ms_per_step"], d["value"], d["roofline"]["frac"], d["failed_cells"]))
except Exception as e:
    print("$m $rx FAILED", e); print(open("gpurun_out/r02w_bench.err").read()[-500:])
P
done; done
