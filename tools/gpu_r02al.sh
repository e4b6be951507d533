#!/bin/bash
# session r02al: full GPU test suite + smoke + default bench line of the final library (all fourteen ERK tables, RK64 mask fix)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee gpurun_out/r02al_pytest_gpu.log
timeout 600 python __graft_entry__.py smoke 2>&1 | tail -3 | tee gpurun_out/r02al_smoke.log
timeout 900 python bench.py --no-cpu-baseline --no-ref-gpu --strong none 2>gpurun_out/r02al_bench.err | tee gpurun_out/r02al_bench_lidryer_cvode_128.json | cut -c1-200
timeout 900 python bench.py --reactor ReactorRK64 --no-cpu-baseline --no-ref-gpu --strong none 2>>gpurun_out/r02al_bench.err | tee gpurun_out/r02al_bench_lidryer_rk64_128.json | cut -c1-200
