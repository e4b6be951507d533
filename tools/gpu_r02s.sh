#!/bin/bash
# session r02s: PeleC application drop-in (CPU and CUDA builds), e2e with thin edge slabs
mkdir -p gpurun_out
python -m pytest tests/test_pelec_dropin.py -m gpu -q -s 2>&1 | tail -30 > gpurun_out/r02s_pytest_pelec.log
cat gpurun_out/r02s_pytest_pelec.log
for e in 0 4 8 16; do
  PRX_HOST_EDGE_PLANES=$e python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-ref-gpu --strong none > gpurun_out/r02s_bench_edge$e.json 2> gpurun_out/r02s_edge$e.err
  python - <<P
import json
try:
    d=json.loads(open("gpurun_out/r02s_bench_edge$e.json").read().strip().splitlines()[-1])
    print("edge $e: device %.2f ms, e2e %.2f ms (%.2f M cells/s)"%(d["ms_per_step"], d["e2e"]["ms_per_step"], d["e2e"]["value"]/1e6))
except Exception as ex:
    print("edge $e FAILED", ex); print(open("gpurun_out/r02s_edge$e.err").read()[-300:])
P
done
PRX_HOST_EDGE_PLANES=8 PRX_HOST_SLABS=3 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-ref-gpu --strong none 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('edge 8 slabs 3: e2e %.2f ms'%d['e2e']['ms_per_step'])"
