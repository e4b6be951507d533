#!/bin/bash
# Round 2, session M: lean one-warp-per-matrix LU / solve + pipelined production-rate gather (dodecane_lu): parity, timing, capture
set -u
TAG=r02m
mkdir -p gpurun_out
O=pelec_b200/lib/obj
timeout 1500 python -m pytest tests -m gpu -q -k "dodecane or coop or every_bdf or state or dropin" 2>&1 | tail -6 | tee gpurun_out/${TAG}_pytest_gpu.log
timeout 600 python bench.py --no-cpu-baseline --mech dodecane_lu --n 32 --steps 2 --warmup 1 2>>gpurun_out/${TAG}_bench.err > gpurun_out/${TAG}_bench_dodecane_32.json
timeout 600 python bench.py --no-cpu-baseline --mech dodecane_lu --n 32 --steps 2 --warmup 1 --rtol 1e-6 --max-order 2 2>>gpurun_out/${TAG}_bench.err > gpurun_out/${TAG}_bench_dodecane_32_pelec_tols.json
PRX_BDF_VARIANT=32 timeout 300 python bench.py --no-cpu-baseline --mech drm19 --n 32 --steps 3 2>>gpurun_out/${TAG}_bench.err > gpurun_out/${TAG}_bench_drm19_coop32_32.json
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/${TAG}_bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], "%.3e %s" % (d["value"], d["unit"]), d.get("ms_per_step"), (d.get("roofline") or {}).get("frac"), d.get("failed_cells"), d.get("counters_per_cell"))
    except Exception as e:
        print(f, "unreadable", e)
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:bdfc_kernel -s 1 -c 1 -f -o gpurun_out/${TAG}_prof_bdfc32w4_dodecane_16 \
   python bench.py --steps 1 --warmup 1 --n 16 --mech dodecane_lu --no-cpu-baseline > gpurun_out/${TAG}_ncu2.log 2>&1
tools/ncu_extract.sh gpurun_out/${TAG}_prof_bdfc32w4_dodecane_16.ncu-rep $O/kernels_dodecane_lu.o bdfc_kernel_dodecane_luILi32ELi4E
grep -v pcsamp gpurun_out/${TAG}_prof_bdfc32w4_dodecane_16_summary.txt | head -20
