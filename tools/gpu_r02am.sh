#!/bin/bash
# session r02am (2 GPUs): the driver's scaling invocation at N = 2 on the final library (weak headline + sharded-domain block, e2e with the new staging pipeline)
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 2>gpurun_out/r02am_bench_n2.err > gpurun_out/r02am_bench_n2.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 2>>gpurun_out/r02am_bench_n2.err > gpurun_out/r02am_bench_reference_n2.json
python - <<PY
import json
for f in ("gpurun_out/r02am_bench_n2.json", "gpurun_out/r02am_bench_reference_n2.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], "%.3e %s" % (d["value"], d["unit"]), d.get("ms_per_step"), "e2e", (d.get("e2e") or {}).get("value"), (d.get("e2e") or {}).get("ms_per_step"), "failed", d.get("failed_cells"))
        for s in d.get("strong", []):
            print("   strong:", s["mechanism"], s["cells"], "boxes", s["boxes"], "gpus", s["n_gpus"], "ms", s["ms_per_step"], "%.3e cells/s" % s["value"])
    except Exception as e:
        print(f, "unreadable", e)
PY
tail -3 gpurun_out/r02am_bench_n2.err
