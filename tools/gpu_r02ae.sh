#!/bin/bash
# session r02ae: host staging pipeline with alternating react streams (PRX_HOST_REACT_STREAMS) x slab count
mkdir -p gpurun_out
PRX_HOST_SLABS=8 PRX_HOST_REACT_STREAMS=3 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "full_size or threads" 2>&1 | tail -5 > gpurun_out/r02ae_pytest_host.log
cat gpurun_out/r02ae_pytest_host.log
for cfg in "1 2" "2 2" "2 4" "2 8" "3 6" "2 12" "4 8" "1 4"; do
  set -- $cfg
  PRX_HOST_REACT_STREAMS=$1 PRX_HOST_SLABS=$2 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-ref-gpu --strong none > gpurun_out/r02ae_bench_rs$1_slabs$2.json 2> gpurun_out/r02ae_rs$1_slabs$2.err
  python - <<P
import json
try:
    d=json.loads(open("gpurun_out/r02ae_bench_rs$1_slabs$2.json").read().strip().splitlines()[-1])
    print("streams $1 slabs $2: device %.2f ms, e2e %.2f ms (%.2f M cells/s) check %s"%(d["ms_per_step"], d["e2e"]["ms_per_step"], d["e2e"]["value"]/1e6, d.get("result_check")))
except Exception as ex:
    print("streams $1 slabs $2 FAILED", ex); print(open("gpurun_out/r02ae_rs$1_slabs$2.err").read()[-300:])
P
done
