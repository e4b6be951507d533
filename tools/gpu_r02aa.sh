#!/bin/bash
# session r02aa: full GPU suite on the final library; default bench
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/r02aa_pytest_gpu.log; cat gpurun_out/r02aa_pytest_gpu.log
python bench.py > gpurun_out/r02aa_bench_default.json 2> gpurun_out/r02aa_bench.err; python - <<'P'
import json
d=json.loads(open("gpurun_out/r02aa_bench_default.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step")}, "e2e", d["e2e"]["value"], "frac", d["roofline"]["frac"], "refgpu", d.get("ref_gpu_baseline",{}).get("value"), "cpu", d.get("cpu_baseline",{}).get("value"), d.get("result_check"))
P
