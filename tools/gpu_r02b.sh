#!/bin/bash
# Round 2, session B: fat-thread kernel (variant 5) - parity, timing at NT = 128/96/64, ncu capture
set -u
TAG=r02b
mkdir -p gpurun_out
O=pelec_b200/lib/obj
timeout 600 python -m pytest tests -m gpu -q -k "every_bdf_kernel_generation and LiDryer" 2>&1 | tail -4 | tee gpurun_out/${TAG}_pytest_gpu.log
for bw in 4 3 2; do
  PRX_BDF_VARIANT=5 PRX_BLOCK_WARPS=$bw timeout 300 python bench.py --no-cpu-baseline 2>>gpurun_out/${TAG}_bench.err > gpurun_out/${TAG}_bench_lidryer_v5_bw$bw.json
done
PRX_BDF_VARIANT=5 timeout 300 python bench.py --no-cpu-baseline --states pmf 2>>gpurun_out/${TAG}_bench.err > gpurun_out/${TAG}_bench_lidryer_v5_pmf.json
PRX_BDF_VARIANT=5 timeout 300 python bench.py --no-cpu-baseline --rtol 1e-6 --atol 1e-10 --max-order 2 2>>gpurun_out/${TAG}_bench.err > gpurun_out/${TAG}_bench_lidryer_v5_pelec_tols.json
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/${TAG}_bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], "%.3e %s" % (d["value"], d["unit"]), d.get("ms_per_step"), (d.get("roofline") or {}).get("frac"), (d.get("e2e") or {}).get("value"), d.get("failed_cells"))
    except Exception as e:
        print(f, "unreadable", e)
PY
tail -3 gpurun_out/${TAG}_bench.err
PRX_BDF_VARIANT=5 timeout 900 ncu --set full --clock-control none --import-source on -k regex:bdf4_kernel -s 1 -c 1 -f -o gpurun_out/${TAG}_prof_bdf4_lidryer_128 \
   python bench.py --steps 1 --warmup 3 --n 128 --no-cpu-baseline > gpurun_out/${TAG}_ncu1.log 2>&1
tools/ncu_extract.sh gpurun_out/${TAG}_prof_bdf4_lidryer_128.ncu-rep $O/kernels_LiDryer.o bdf4_kernel
head -24 gpurun_out/${TAG}_prof_bdf4_lidryer_128_summary.txt
