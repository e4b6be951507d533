#!/bin/bash
# session r02ah: slab-pipeline equality test, ReactorArkode 1-D react, default bench of the rebuilt library
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_arkode.py -m gpu -q -s -k "slab_pipeline or arkode" 2>&1 | tail -12 | tee gpurun_out/r02ah_pytest_new.log
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-ref-gpu --strong none > gpurun_out/r02ah_bench_default.json 2> gpurun_out/r02ah_bench.err
python - <<P
import json
d=json.loads(open("gpurun_out/r02ah_bench_default.json").read().strip().splitlines()[-1])
print("default: %.2f ms, e2e %.2f ms (%.2f M cells/s), failed %s, check %s"%(d["ms_per_step"], d["e2e"]["ms_per_step"], d["e2e"]["value"]/1e6, d["failed_cells"], d.get("result_check")))
P
