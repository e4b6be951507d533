#!/bin/bash
# session r02x: shipped-input PMF drop-in test; extra-mechanism throughput; ncu of the ctoprim / transport kernels
mkdir -p gpurun_out
TAG=r02x
O=pelec_b200/lib/obj
python -m pytest tests/test_pelec_dropin.py -m gpu -q -s -k shipped 2>&1 | tail -6 > gpurun_out/${TAG}_pytest_shipped.log; cat gpurun_out/${TAG}_pytest_shipped.log
for m in BurkeDryer FFCM1_Red; do for rx in ReactorCvode ReactorRK64; do
  python bench.py --mech $m --n 64 --reactor $rx --steps 3 --warmup 3 --no-cpu-baseline --no-ref-gpu --strong none > gpurun_out/${TAG}_bench_${m}_${rx}_64.json 2> gpurun_out/${TAG}_bench.err
  python - <<P
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_bench_${m}_${rx}_64.json").read().strip().splitlines()[-1])
    print("$m $rx 64^3: %.2f ms, %.3g cells/s, roofline %.3f, failed %s"%(d["ms_per_step"], d["value"], d["roofline"]["frac"], d["failed_cells"]))
except Exception as e:
    print("$m $rx FAILED", e); print(open("gpurun_out/${TAG}_bench.err").read()[-500:])
P
done; done
for k in ctoprim transport; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:${k}_kernel -s 3 -c 1 -f -o gpurun_out/${TAG}_prof_${k}_lidryer_192 \
     python tools/bench_eos.py --mech LiDryer --n 192 --kernel $k > gpurun_out/${TAG}_ncu_$k.log 2>&1
  tools/ncu_extract.sh gpurun_out/${TAG}_prof_${k}_lidryer_192.ncu-rep $O/kernels_LiDryer.o ${k}_kernel
  rm -f gpurun_out/${TAG}_prof_${k}_lidryer_192_source.csv.gz
  head -30 gpurun_out/${TAG}_prof_${k}_lidryer_192_summary.txt | cut -c1-160
done
