#!/bin/bash
# Round 2, session N: dodecane_lu cooperative kernel knobs - warps per block 1/2/4/8, one or two block barriers, LU batch 8, 20 warps/SM at 96 registers;
# default bench with the result check
set -u
TAG=r02n
mkdir -p gpurun_out
D="--no-cpu-baseline --mech dodecane_lu --n 32 --steps 2 --warmup 1"
for w in 1 2 4 8; do
  PRX_BLOCK_WARPS=$w timeout 600 python bench.py $D 2>>gpurun_out/${TAG}_bench.err > gpurun_out/${TAG}_bench_dodecane_32_w$w.json
  PRX_BLOCK_WARPS=$w PRX_COOP_BARRIERS=1 timeout 600 python bench.py $D 2>>gpurun_out/${TAG}_bench.err > gpurun_out/${TAG}_bench_dodecane_32_w${w}_b1.json
done
PRX_LIB_PATH=$PWD/pelec_b200/lib/libpelerx_b200_cb8.so timeout 600 python bench.py $D 2>>gpurun_out/${TAG}_bench.err > gpurun_out/${TAG}_bench_dodecane_32_cb8.json
PRX_LIB_PATH=$PWD/pelec_b200/lib/libpelerx_b200_mb20.so timeout 600 python bench.py $D 2>>gpurun_out/${TAG}_bench.err > gpurun_out/${TAG}_bench_dodecane_32_mb20.json
timeout 600 python bench.py --cpu-seconds 10 --strong none 2>>gpurun_out/${TAG}_bench.err > gpurun_out/${TAG}_bench_lidryer_cvode_128.json
timeout 600 python bench.py --cpu-seconds 10 --reactor ReactorRK64 2>>gpurun_out/${TAG}_bench.err > gpurun_out/${TAG}_bench_lidryer_rk64_128.json
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/${TAG}_bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], "%.3e %s" % (d["value"], d["unit"]), d.get("ms_per_step"), (d.get("roofline") or {}).get("frac"), d.get("failed_cells"), d.get("result_check"))
    except Exception as e:
        print(f, "unreadable", e)
PY
tail -3 gpurun_out/${TAG}_bench.err
