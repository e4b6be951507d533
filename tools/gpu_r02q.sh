#!/bin/bash
# session r02q: new drop-in tests (PeleC application, CUDA build of the reference), full GPU suite, bench with ref_gpu_baseline
mkdir -p gpurun_out
python -m pytest tests/test_react_state_golden.py tests/test_dropin_gpu_build.py tests/test_pelec_dropin.py -m gpu -x -q -s 2>&1 | tail -25 > gpurun_out/r02q_pytest_new.log
cat gpurun_out/r02q_pytest_new.log
python bench.py --steps 5 --warmup 3 > gpurun_out/r02q_bench_default.json 2> gpurun_out/r02q_bench.err
python - <<'P'
import json
d=json.loads(open("gpurun_out/r02q_bench_default.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step")}, d["e2e"]["value"], d["roofline"]["frac"], d.get("ref_gpu_baseline"), d.get("cpu_baseline",{}).get("value"))
P
for r in ReactorCvodeB200 ReactorRK64B200; do python tools/ref_gpu_bench.py --mech LiDryer --n 128 --reactor $r --reps 3 --check > gpurun_out/r02q_shimgpu_$r.json 2>gpurun_out/r02q_shimgpu_$r.err; tail -c 700 gpurun_out/r02q_shimgpu_$r.json; done
python tools/ref_gpu_bench.py --mech drm19 --n 64 --solve-type GMRES --reps 2 --check > gpurun_out/r02q_refgpu_drm19_gmres_64.json 2>gpurun_out/r02q_refgpu_drm19.err; tail -c 700 gpurun_out/r02q_refgpu_drm19_gmres_64.json
