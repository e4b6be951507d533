#!/bin/bash
# Round 2, session I: new tests, sanitizer, EOS kernel bandwidth, final captures of the default kernels (traffic on the timed size)
set -u
TAG=r02i
mkdir -p gpurun_out
O=pelec_b200/lib/obj
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee gpurun_out/${TAG}_pytest_gpu.log
bash tools/gpu_sanitizer.sh ${TAG}
timeout 300 python tools/bench_eos.py > gpurun_out/${TAG}_eos_lidryer_256.json 2>gpurun_out/${TAG}_eos.err; cat gpurun_out/${TAG}_eos_lidryer_256.json
timeout 300 python tools/bench_eos.py --mech drm19 > gpurun_out/${TAG}_eos_drm19_256.json 2>>gpurun_out/${TAG}_eos.err; cat gpurun_out/${TAG}_eos_drm19_256.json
timeout 900 ncu --set full --clock-control none --import-source on -k regex:bdf_kernel -s 1 -c 1 -f -o gpurun_out/${TAG}_prof_bdf_lidryer_128 \
   python bench.py --steps 1 --warmup 3 --n 128 --no-cpu-baseline --strong none > gpurun_out/${TAG}_ncu1.log 2>&1
tools/ncu_extract.sh gpurun_out/${TAG}_prof_bdf_lidryer_128.ncu-rep $O/kernels_LiDryer.o bdf_kernel_LiDryer
timeout 900 ncu --set full --clock-control none --import-source on -k regex:rk64_kernel -s 1 -c 1 -f -o gpurun_out/${TAG}_prof_rk64_lidryer_128 \
   python bench.py --steps 1 --warmup 3 --n 128 --reactor ReactorRK64 --no-cpu-baseline > gpurun_out/${TAG}_ncu2.log 2>&1
tools/ncu_extract.sh gpurun_out/${TAG}_prof_rk64_lidryer_128.ncu-rep $O/kernels_LiDryer.o rk64_kernel
timeout 900 ncu --set full --clock-control none --import-source on -k regex:bdf_kernel -s 1 -c 1 -f -o gpurun_out/${TAG}_prof_bdf_drm19_64 \
   python bench.py --steps 1 --warmup 3 --n 64 --mech drm19 --no-cpu-baseline > gpurun_out/${TAG}_ncu3.log 2>&1
tools/ncu_extract.sh gpurun_out/${TAG}_prof_bdf_drm19_64.ncu-rep $O/kernels_drm19.o bdf_kernel_drm19
timeout 900 ncu --set full --clock-control none --import-source on -k regex:compute_temp -s 1 -c 1 -f -o gpurun_out/${TAG}_prof_eos_lidryer_256 \
   python tools/bench_eos.py > gpurun_out/${TAG}_ncu4.log 2>&1
tools/ncu_extract.sh gpurun_out/${TAG}_prof_eos_lidryer_256.ncu-rep $O/kernels_LiDryer.o compute_temp_kernel
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/${TAG}_launches_lidryer_128.csv \
   python bench.py --steps 2 --warmup 3 --no-cpu-baseline --strong none > gpurun_out/${TAG}_ncu_launches.log 2>&1
timeout 300 python bench.py --no-cpu-baseline --strong none 2>>gpurun_out/${TAG}_bench.err > gpurun_out/${TAG}_bench_lidryer_cvode_128.json
timeout 300 python bench.py --no-cpu-baseline --reactor ReactorRK64 2>>gpurun_out/${TAG}_bench.err > gpurun_out/${TAG}_bench_lidryer_rk64_128.json
timeout 300 python bench.py --no-cpu-baseline --mech drm19 --n 64 2>>gpurun_out/${TAG}_bench.err > gpurun_out/${TAG}_bench_drm19_cvode_64.json
timeout 300 python bench.py --no-cpu-baseline --mech drm19 --n 128 --steps 3 2>>gpurun_out/${TAG}_bench.err > gpurun_out/${TAG}_bench_drm19_cvode_128.json
timeout 300 python bench.py --no-cpu-baseline --mech drm19 --n 32 --reactor ReactorRK64 2>>gpurun_out/${TAG}_bench.err > gpurun_out/${TAG}_bench_drm19_rk64_32.json
timeout 300 python bench.py --no-cpu-baseline --rtol 1e-6 --atol 1e-10 --max-order 2 2>>gpurun_out/${TAG}_bench.err > gpurun_out/${TAG}_bench_lidryer_cvode_128_pelec_tols.json
timeout 300 python bench.py --no-cpu-baseline --states pmf 2>>gpurun_out/${TAG}_bench.err > gpurun_out/${TAG}_bench_lidryer_cvode_128_pmf_states.json
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/${TAG}_bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], "%.3e %s" % (d["value"], d["unit"]), d.get("ms_per_step"), (d.get("roofline") or {}).get("frac"), (d.get("e2e") or {}).get("value"), d.get("failed_cells"))
    except Exception as e:
        print(f, "unreadable", e)
PY
head -4 gpurun_out/${TAG}_prof_eos_lidryer_256_summary.txt
