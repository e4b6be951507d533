#!/bin/bash
# session r02az: BurkeDryer (13 species, class LiDryer) with the 384 x 2 block shape vs its 224 x 3
mkdir -p gpurun_out
L=pelec_b200/lib/libpelerx_b200
PRX_LIB_PATH=${L}_bd384.so python -m pytest tests/test_extra_mechanisms.py -m gpu -q -k BurkeDryer 2>&1 | tail -2 | tee gpurun_out/r02az_pytest_bd384.log
for rep in 1 2; do
for v in base bd384; do
  lib=${L}_$v.so; [ $v = base ] && lib=$L.so
  PRX_LIB_PATH=$lib python bench.py --mech BurkeDryer --n 64 --steps 5 --warmup 3 --no-cpu-baseline --no-ref-gpu --strong none > gpurun_out/r02az_bench_BurkeDryer_${v}_$rep.json 2> gpurun_out/r02az_$v.err
  python - <<P
import json
try:
    d=json.loads(open("gpurun_out/r02az_bench_BurkeDryer_${v}_$rep.json").read().strip().splitlines()[-1])
    print("$v run $rep: %.2f ms, failed %s"%(d["ms_per_step"], d["failed_cells"]))
except Exception as e:
    print("$v FAILED", e); print(open("gpurun_out/r02az_$v.err").read()[-300:])
P
done
done
