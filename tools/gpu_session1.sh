#!/bin/bash
# GPU session 1: parity tests, first bench lines, ncu launch list + full capture of the headline kernel.
set -u
mkdir -p gpurun_out
nvidia-smi > gpurun_out/nvidia_smi.txt 2>&1
echo "== pytest -m gpu" 
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 | tee gpurun_out/pytest_gpu.log
echo "== smoke"
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -5 | tee gpurun_out/smoke.log
echo "== bench LiDryer cvode 128^3"
timeout 600 python bench.py --steps 5 --warmup 3 2>gpurun_out/bench_lidryer.err | tee gpurun_out/bench_lidryer.json
echo "== bench LiDryer rk64 128^3"
timeout 600 python bench.py --steps 3 --warmup 3 --reactor ReactorRK64 --no-cpu-baseline 2>gpurun_out/bench_lidryer_rk64.err | tee gpurun_out/bench_lidryer_rk64.json
echo "== bench drm19 cvode 64^3"
timeout 600 python bench.py --steps 3 --warmup 3 --mech drm19 --n 64 --no-cpu-baseline 2>gpurun_out/bench_drm19.err | tee gpurun_out/bench_drm19.json
echo "== bench dodecane cvode 32^3"
timeout 600 python bench.py --steps 3 --warmup 3 --mech dodecane_lu --n 32 --no-cpu-baseline 2>gpurun_out/bench_dodecane.err | tee gpurun_out/bench_dodecane.json
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_lidryer.csv \
   python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
echo "== ncu full"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:bdf_kernel -s 1 -c 1 -o gpurun_out/prof_bdf_lidryer \
   python bench.py --steps 1 --warmup 3 --n 64 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
