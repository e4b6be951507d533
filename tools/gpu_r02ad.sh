#!/bin/bash
# session r02ad: ReactorArkode: parity, PMF application drop-in, timing beside the reference's own GPU ReactorArkode
mkdir -p gpurun_out
python -m pytest tests/test_arkode.py tests/test_pelec_dropin.py -m gpu -q -s -k "arkode or Arkode" 2>&1 | tail -12 > gpurun_out/r02ad_pytest_arkode.log; cat gpurun_out/r02ad_pytest_arkode.log
for r in ReactorArkode ReactorArkodeB200; do for n in 64 128; do
  python tools/ref_gpu_bench.py --mech LiDryer --n $n --reactor $r --rtol 1e-6 --atol 1e-10 --reps 3 > gpurun_out/r02ad_arkode_${r}_$n.json 2> gpurun_out/r02ad_arkode.err
  python - <<P
import json
try:
    d=json.loads(open("gpurun_out/r02ad_arkode_${r}_$n.json").read().strip().splitlines()[-1])
    print("$r $n^3:", [round(x,2) for x in d["ms_per_call"]], "ms, nfe", d["react_return"], "%.3g cells/s"%d["value"])
except Exception as e:
    print("$r $n FAILED", e); print(open("gpurun_out/r02ad_arkode.err").read()[-400:])
P
done; done
