#!/bin/bash
set -u
for s in 1 2 4 6 8 12; do
PRX_HOST_SLABS=$s timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('slabs $s  device %.1f ms  e2e %.3e cells/s' % (d['ms_per_step'], d['e2e']['value']))"
done
PRX_HOST_SLABS=8 timeout 600 python bench.py --states pmf --steps 3 --warmup 2 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('pmf slabs 8  device %.1f ms  e2e %.3e cells/s' % (d['ms_per_step'], d['e2e']['value']))"
