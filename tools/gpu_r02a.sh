#!/bin/bash
# Round 2, session A: baseline of the round-1 build - parity, SASS opcode counts (flop-model validation), RK64 capture,
# DRAM traffic on the timed launch size, small boxes.
set -u
TAG=r02a
mkdir -p gpurun_out
O=pelec_b200/lib/obj
echo "== pytest -m gpu"
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -4 | tee gpurun_out/${TAG}_pytest_gpu.log
echo "== small boxes"
for n in 16 32 64; do
  timeout 300 python bench.py --n $n --no-cpu-baseline 2>>gpurun_out/${TAG}_bench.err > gpurun_out/${TAG}_bench_lidryer_cvode_$n.json
  timeout 300 python bench.py --n $n --mech drm19 --no-cpu-baseline 2>>gpurun_out/${TAG}_bench.err > gpurun_out/${TAG}_bench_drm19_cvode_$n.json
done
timeout 300 python bench.py --no-cpu-baseline 2>>gpurun_out/${TAG}_bench.err > gpurun_out/${TAG}_bench_lidryer_cvode_128.json
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/${TAG}_bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], "%.3e %s" % (d["value"], d["unit"]), d.get("ms_per_step"), (d.get("roofline") or {}).get("frac"), (d.get("e2e") or {}).get("value"))
    except Exception as e:
        print(f, "unreadable", e)
PY
echo "== ncu full: bdf LiDryer 128 (traffic on the timed launch size)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:bdf_kernel -s 1 -c 1 -f -o gpurun_out/${TAG}_prof_bdf_lidryer_128 \
   python bench.py --steps 1 --warmup 3 --n 128 --no-cpu-baseline > gpurun_out/${TAG}_ncu1.log 2>&1
tools/ncu_extract.sh gpurun_out/${TAG}_prof_bdf_lidryer_128.ncu-rep $O/kernels_LiDryer.o bdf_kernel
echo "== ncu full: rk64 LiDryer 64"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:rk64_kernel -s 1 -c 1 -f -o gpurun_out/${TAG}_prof_rk64_lidryer_64 \
   python bench.py --steps 1 --warmup 3 --n 64 --reactor ReactorRK64 --no-cpu-baseline > gpurun_out/${TAG}_ncu2.log 2>&1
tools/ncu_extract.sh gpurun_out/${TAG}_prof_rk64_lidryer_64.ncu-rep $O/kernels_LiDryer.o rk64_kernel
echo "== ncu full: bdf drm19 64"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:bdf_kernel -s 1 -c 1 -f -o gpurun_out/${TAG}_prof_bdf_drm19_64 \
   python bench.py --steps 1 --warmup 3 --n 64 --mech drm19 --no-cpu-baseline > gpurun_out/${TAG}_ncu3.log 2>&1
tools/ncu_extract.sh gpurun_out/${TAG}_prof_bdf_drm19_64.ncu-rep $O/kernels_drm19.o bdf_kernel
echo "== ncu full: rk64 drm19 32"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:rk64_kernel -s 1 -c 1 -f -o gpurun_out/${TAG}_prof_rk64_drm19_32 \
   python bench.py --steps 1 --warmup 3 --n 32 --mech drm19 --reactor ReactorRK64 --no-cpu-baseline > gpurun_out/${TAG}_ncu4.log 2>&1
tools/ncu_extract.sh gpurun_out/${TAG}_prof_rk64_drm19_32.ncu-rep $O/kernels_drm19.o rk64_kernel
ls -la gpurun_out | tail -30
