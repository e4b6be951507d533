#!/bin/bash
# session r02v: ctoprim / transport kernels: parity + throughput; full GPU suite
mkdir -p gpurun_out
python -m pytest tests/test_eos_transport.py -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r02v_pytest_eos.log; cat gpurun_out/r02v_pytest_eos.log
for k in ctoprim transport; do for m in LiDryer drm19; do
  python tools/bench_eos.py --mech $m --n 192 --kernel $k > gpurun_out/r02v_eos_${k}_${m}_192.json 2>gpurun_out/r02v_eos.err; tail -c 400 gpurun_out/r02v_eos_${k}_${m}_192.json
done; done
python tools/bench_eos.py --mech dodecane_lu --n 96 --kernel transport > gpurun_out/r02v_eos_transport_dodecane_lu_96.json 2>>gpurun_out/r02v_eos.err; tail -c 400 gpurun_out/r02v_eos_transport_dodecane_lu_96.json
python tools/bench_eos.py --mech dodecane_lu --n 96 --kernel ctoprim > gpurun_out/r02v_eos_ctoprim_dodecane_lu_96.json 2>>gpurun_out/r02v_eos.err; tail -c 400 gpurun_out/r02v_eos_ctoprim_dodecane_lu_96.json
python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r02v_pytest_gpu.log; cat gpurun_out/r02v_pytest_gpu.log
