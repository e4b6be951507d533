#!/bin/bash
# Round 2, session C: literal pool + custom exp/log10 build - parity and timing; cooperative-kernel baselines with by-line captures
set -u
TAG=r02c
mkdir -p gpurun_out
O=pelec_b200/lib/obj
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee gpurun_out/${TAG}_pytest_gpu.log
timeout 300 python bench.py --no-cpu-baseline 2>>gpurun_out/${TAG}_bench.err > gpurun_out/${TAG}_bench_lidryer_cvode_128.json
timeout 300 python bench.py --no-cpu-baseline --reactor ReactorRK64 2>>gpurun_out/${TAG}_bench.err > gpurun_out/${TAG}_bench_lidryer_rk64_128.json
timeout 300 python bench.py --no-cpu-baseline --mech drm19 --n 64 2>>gpurun_out/${TAG}_bench.err > gpurun_out/${TAG}_bench_drm19_cvode_64.json
timeout 300 python bench.py --no-cpu-baseline --mech drm19 --n 32 --reactor ReactorRK64 2>>gpurun_out/${TAG}_bench.err > gpurun_out/${TAG}_bench_drm19_rk64_32.json
PRX_BDF_VARIANT=8 timeout 300 python bench.py --no-cpu-baseline --n 64 --steps 3 2>>gpurun_out/${TAG}_bench.err > gpurun_out/${TAG}_bench_lidryer_coop8_64.json
PRX_BDF_VARIANT=8 timeout 300 python bench.py --no-cpu-baseline --mech drm19 --n 32 --steps 3 2>>gpurun_out/${TAG}_bench.err > gpurun_out/${TAG}_bench_drm19_coop8_32.json
PRX_BDF_VARIANT=16 timeout 300 python bench.py --no-cpu-baseline --mech drm19 --n 32 --steps 3 2>>gpurun_out/${TAG}_bench.err > gpurun_out/${TAG}_bench_drm19_coop16_32.json
timeout 600 python bench.py --no-cpu-baseline --mech dodecane_lu --n 32 --steps 2 --warmup 1 2>>gpurun_out/${TAG}_bench.err > gpurun_out/${TAG}_bench_dodecane_cvode_32.json
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
PRX_BDF_VARIANT=8 timeout 900 ncu --set full --clock-control none --import-source on -k regex:bdfc_kernel -s 1 -c 1 -f -o gpurun_out/${TAG}_prof_bdfc8_lidryer_64 \
   python bench.py --steps 1 --warmup 3 --n 64 --no-cpu-baseline > gpurun_out/${TAG}_ncu1.log 2>&1
tools/ncu_extract.sh gpurun_out/${TAG}_prof_bdfc8_lidryer_64.ncu-rep $O/kernels_LiDryer.o bdfc_kernel_LiDryerILi8E
timeout 900 ncu --set full --clock-control none --import-source on -k regex:bdfc_kernel -s 1 -c 1 -f -o gpurun_out/${TAG}_prof_bdfc32_dodecane_16 \
   python bench.py --steps 1 --warmup 1 --n 16 --mech dodecane_lu --no-cpu-baseline > gpurun_out/${TAG}_ncu2.log 2>&1
tools/ncu_extract.sh gpurun_out/${TAG}_prof_bdfc32_dodecane_16.ncu-rep $O/kernels_dodecane_lu.o bdfc_kernel_dodecane_luILi32E
timeout 900 ncu --set full --clock-control none --import-source on -k regex:rk64_kernel -s 1 -c 1 -f -o gpurun_out/${TAG}_prof_rk64_lidryer_64 \
   python bench.py --steps 1 --warmup 3 --n 64 --reactor ReactorRK64 --no-cpu-baseline > gpurun_out/${TAG}_ncu3.log 2>&1
tools/ncu_extract.sh gpurun_out/${TAG}_prof_rk64_lidryer_64.ncu-rep $O/kernels_LiDryer.o rk64_kernel
head -20 gpurun_out/${TAG}_prof_rk64_lidryer_64_summary.txt
