#!/bin/bash
# Final GPU session of the round: parity, smoke, the bench lines that go to profiles/, launch list and ncu capture of the
# default kernel.  usage: tools/gpu_final.sh <tag>
set -u
TAG=${1:-r01i}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > gpurun_out/${TAG}_nvidia_smi.txt 2>&1
echo "== pytest -m gpu"
timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee gpurun_out/${TAG}_pytest_gpu.log
echo "== smoke"
timeout 600 python __graft_entry__.py smoke 2>&1 | tail -3 | tee gpurun_out/${TAG}_smoke.log
echo "== bench (default: 128^3 LiDryer ReactorCvode, with cpu_baseline)"
timeout 900 python bench.py 2>gpurun_out/${TAG}_bench.err | tee gpurun_out/${TAG}_bench_lidryer_cvode_128.json | cut -c1-400
echo "== bench --impl reference"
timeout 900 python bench.py --impl reference 2>>gpurun_out/${TAG}_bench.err | tee gpurun_out/${TAG}_bench_reference_lidryer_cvode_128.json | cut -c1-300
echo "== other configurations"
timeout 600 python bench.py --reactor ReactorRK64 --no-cpu-baseline 2>>gpurun_out/${TAG}_bench.err > gpurun_out/${TAG}_bench_lidryer_rk64_128.json
timeout 600 python bench.py --mech drm19 --n 64 --no-cpu-baseline 2>>gpurun_out/${TAG}_bench.err > gpurun_out/${TAG}_bench_drm19_cvode_64.json
timeout 900 python bench.py --mech drm19 --n 128 --steps 3 --no-cpu-baseline 2>>gpurun_out/${TAG}_bench.err > gpurun_out/${TAG}_bench_drm19_cvode_128.json
timeout 900 python bench.py --mech dodecane_lu --n 32 --steps 3 --no-cpu-baseline 2>>gpurun_out/${TAG}_bench.err > gpurun_out/${TAG}_bench_dodecane_cvode_32.json
timeout 600 python bench.py --rtol 1e-6 --atol 1e-10 --max-order 2 --no-cpu-baseline 2>>gpurun_out/${TAG}_bench.err > gpurun_out/${TAG}_bench_lidryer_cvode_128_pelec_tols.json
PRX_SORT_CELLS=0 timeout 600 python bench.py --steps 3 --no-cpu-baseline 2>>gpurun_out/${TAG}_bench.err > gpurun_out/${TAG}_bench_x_lidryer_unsorted.json
timeout 900 python bench.py --states pmf --cpu-seconds 10 2>>gpurun_out/${TAG}_bench.err > gpurun_out/${TAG}_bench_lidryer_cvode_128_pmf_states.json
PRX_SORT_CELLS=0 timeout 600 python bench.py --states pmf --steps 3 --no-cpu-baseline 2>>gpurun_out/${TAG}_bench.err > gpurun_out/${TAG}_bench_x_lidryer_pmf_states_unsorted.json
timeout 600 python bench.py --states pmf --mech drm19 --n 64 --no-cpu-baseline 2>>gpurun_out/${TAG}_bench.err > gpurun_out/${TAG}_bench_drm19_cvode_64_pmf_states.json
timeout 600 python bench.py --states pmf --reactor ReactorRK64 --no-cpu-baseline 2>>gpurun_out/${TAG}_bench.err > gpurun_out/${TAG}_bench_lidryer_rk64_128_pmf_states.json
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/${TAG}_bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], "%.3e %s" % (d["value"], d["unit"]), d.get("ms_per_step"), (d.get("roofline") or {}).get("frac"), (d.get("e2e") or {}).get("value"))
    except Exception as e:
        print(f, "unreadable", e)
PY
tail -5 gpurun_out/${TAG}_bench.err
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/${TAG}_launches_lidryer_128.csv \
   python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_launches.log 2>&1
echo "== ncu full, default kernel"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:bdf_kernel -s 1 -c 1 -f -o gpurun_out/${TAG}_prof_bdf_lidryer \
   python bench.py --steps 1 --warmup 3 --n 64 --no-cpu-baseline > gpurun_out/${TAG}_ncu_full.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_full.log | cut -c1-200
