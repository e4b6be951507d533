#!/bin/bash
# Round 2, session H (2 GPUs): full GPU suite, 2-rank bench with the sharded-domain (strong) block, dodecane W=8, coop captures
set -u
TAG=r02h
mkdir -p gpurun_out
O=pelec_b200/lib/obj
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -8 | tee gpurun_out/${TAG}_pytest_gpu.log
timeout 900 python bench.py --no-cpu-baseline 2>gpurun_out/${TAG}_bench_n1.err > gpurun_out/${TAG}_bench_n1.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --no-cpu-baseline 2>gpurun_out/${TAG}_bench_n2.err > gpurun_out/${TAG}_bench_n2.json
PRX_BLOCK_WARPS=8 timeout 600 python bench.py --no-cpu-baseline --mech dodecane_lu --n 32 --steps 2 --warmup 1 2>>gpurun_out/${TAG}_bench.err > gpurun_out/${TAG}_bench_dodecane_32_w8.json
timeout 600 python bench.py --no-cpu-baseline --mech dodecane_lu --n 32 --steps 2 --warmup 1 2>>gpurun_out/${TAG}_bench.err > gpurun_out/${TAG}_bench_dodecane_32_w4.json
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/${TAG}_bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], "%.3e %s" % (d["value"], d["unit"]), d.get("ms_per_step"), (d.get("roofline") or {}).get("frac"), (d.get("e2e") or {}).get("value"), d.get("failed_cells"))
        for s in d.get("strong", []):
            print("   strong:", s["mechanism"], s["cells"], "boxes", s["boxes"], "gpus", s["n_gpus"], "ms", s["ms_per_step"], "%.3e cells/s" % s["value"], "imb cost %.3f ms %.3f" % (s["imbalance_cost"], s["imbalance_ms"]))
    except Exception as e:
        print(f, "unreadable", e)
PY
tail -3 gpurun_out/${TAG}_bench_n2.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:bdfc_kernel -s 1 -c 1 -f -o gpurun_out/${TAG}_prof_bdfc32w4_dodecane_16 \
   python bench.py --steps 1 --warmup 1 --n 16 --mech dodecane_lu --no-cpu-baseline > gpurun_out/${TAG}_ncu2.log 2>&1
tools/ncu_extract.sh gpurun_out/${TAG}_prof_bdfc32w4_dodecane_16.ncu-rep $O/kernels_dodecane_lu.o bdfc_kernel_dodecane_luILi32ELi4E
head -20 gpurun_out/${TAG}_prof_bdfc32w4_dodecane_16_summary.txt
