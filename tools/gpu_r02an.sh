#!/bin/bash
# session r02an (8 GPUs): weak-scaling headline + e2e through the host staging pipeline with eight ranks staging at once
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu-baseline --no-ref-gpu --strong none 2>gpurun_out/r02an_bench_n8.err > gpurun_out/r02an_bench_n8.json
python - <<PY
import json
d = json.loads(open("gpurun_out/r02an_bench_n8.json").read().strip().splitlines()[-1])
print("n8: %.3e cells/s %.2f ms, e2e %.3e cells/s %.2f ms, pcie %.2f GB/s per rank" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"].get("pcie_GBps_per_rank", 0)))
PY
tail -2 gpurun_out/r02an_bench_n8.err
