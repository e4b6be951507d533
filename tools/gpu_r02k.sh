#!/bin/bash
# Round 2, session K (8 GPUs): the driver's scaling invocation - weak 128^3 LiDryer per rank + sharded-domain blocks
# (256^3 drm19 over 8 ranks, 128^3 dodecane_lu over 8 ranks)
set -u
TAG=r02k
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,clocks.max.sm --format=csv > gpurun_out/${TAG}_nvidia_smi.txt 2>&1
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 2>gpurun_out/${TAG}_bench_n8.err > gpurun_out/${TAG}_bench_n8.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 4 --no-cpu-baseline 2>gpurun_out/${TAG}_bench_n4.err > gpurun_out/${TAG}_bench_n4.json
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/${TAG}_bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], "%.3e %s" % (d["value"], d["unit"]), d.get("ms_per_step"), (d.get("roofline") or {}).get("frac"), d.get("e2e"), d.get("failed_cells"), d.get("cpu_baseline"))
        for s in d.get("strong", []):
            print("   strong:", s["mechanism"], s["cells"], "boxes", s["boxes"], "gpus", s["n_gpus"], "ms", s["ms_per_step"], "%.3e cells/s" % s["value"], "imb cost %.3f ms %.3f" % (s["imbalance_cost"], s["imbalance_ms"]), "failed", s["failed_cells"])
    except Exception as e:
        print(f, "unreadable", e)
PY
tail -4 gpurun_out/${TAG}_bench_n8.err
