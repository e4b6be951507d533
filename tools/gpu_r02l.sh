#!/bin/bash
# Round 2, session L (8 GPUs): slab count of the host staging pipeline under 8 concurrent ranks (e2e efficiency)
set -u
TAG=r02l
mkdir -p gpurun_out
for s in 2 3 4 6 8; do
  PRX_HOST_SLABS=$s timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2952$s bench.py --gpus 8 --steps 3 --no-cpu-baseline --strong none 2>gpurun_out/${TAG}_n8_slabs$s.err > gpurun_out/${TAG}_bench_n8_slabs$s.json
done
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/${TAG}_bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], "%.4e" % d["value"], d.get("ms_per_step"), "e2e %.4e" % d["e2e"]["value"], "e2e ms", d["e2e"]["ms_per_step"], "pcie", d["e2e"]["pcie_GBps_per_rank"])
    except Exception as e:
        print(f, "unreadable", e)
PY
