#!/bin/bash
# session r02au: evidence of the round's final build - parity, smoke, the bench lines, launch list of a step
set -u
TAG=r02au
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > gpurun_out/${TAG}_nvidia_smi.txt 2>&1
echo "== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee gpurun_out/${TAG}_pytest_gpu.log
echo "== smoke"
timeout 600 python __graft_entry__.py smoke 2>&1 | tail -3 | tee gpurun_out/${TAG}_smoke.log
echo "== bench (default: 128^3 LiDryer ReactorCvode, with cpu_baseline, strong block, reference GPU baseline)"
timeout 900 python bench.py 2>gpurun_out/${TAG}_bench.err | tee gpurun_out/${TAG}_bench_lidryer_cvode_128.json | cut -c1-300
echo "== bench --impl reference"
timeout 900 python bench.py --impl reference 2>>gpurun_out/${TAG}_bench.err | tee gpurun_out/${TAG}_bench_reference_lidryer_cvode_128.json | cut -c1-300
echo "== other configurations"
X="--no-cpu-baseline --no-ref-gpu --strong none"
timeout 600 python bench.py --reactor ReactorRK64 $X 2>>gpurun_out/${TAG}_bench.err > gpurun_out/${TAG}_bench_lidryer_rk64_128.json
timeout 600 python bench.py --mech drm19 --n 64 $X 2>>gpurun_out/${TAG}_bench.err > gpurun_out/${TAG}_bench_drm19_cvode_64.json
timeout 600 python bench.py --rtol 1e-6 --atol 1e-10 --max-order 2 $X 2>>gpurun_out/${TAG}_bench.err > gpurun_out/${TAG}_bench_lidryer_cvode_128_pelec_tols.json
timeout 600 python bench.py --states pmf $X 2>>gpurun_out/${TAG}_bench.err > gpurun_out/${TAG}_bench_lidryer_cvode_128_pmf_states.json
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/${TAG}_bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], "%.3e %s" % (d["value"], d["unit"]), "%.2f ms" % d.get("ms_per_step", 0), "frac", (d.get("roofline") or {}).get("frac"), "e2e", (d.get("e2e") or {}).get("value"), (d.get("e2e") or {}).get("ms_per_step"))
    except Exception as e:
        print(f, "unreadable", e)
PY
tail -5 gpurun_out/${TAG}_bench.err
