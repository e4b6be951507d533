#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 900 python bench.py --states pmf --cpu-seconds 10 2>gpurun_out/pmf.err > gpurun_out/r01j_bench_lidryer_cvode_128_pmf_states.json
timeout 900 python bench.py --states pmf --rtol 1e-6 --atol 1e-10 --max-order 2 --no-cpu-baseline 2>>gpurun_out/pmf.err > gpurun_out/r01j_bench_lidryer_cvode_128_pmf_states_pelec_tols.json
timeout 900 python bench.py --states pmf --mech drm19 --n 64 --no-cpu-baseline 2>>gpurun_out/pmf.err > gpurun_out/r01j_bench_drm19_cvode_64_pmf_states.json
timeout 900 python bench.py --states pmf --reactor ReactorRK64 --no-cpu-baseline 2>>gpurun_out/pmf.err > gpurun_out/r01j_bench_lidryer_rk64_128_pmf_states.json
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/r01j_bench_*pmf*.json")):
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print(f.split("/")[-1], "%.3e %s" % (d["value"], d["unit"]), "%.1f ms" % d["ms_per_step"], "frac %.4f" % d["roofline"]["frac"], "e2e %.3e" % d["e2e"]["value"], d["counters_per_cell"], (d.get("cpu_baseline") or {}).get("value"))
PY
tail -3 gpurun_out/pmf.err
