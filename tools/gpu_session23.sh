#!/bin/bash
set -u
for w in 0 8; do
PRX_MAX_WARPS_PER_SM=$w timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('lidryer max_warps $w  %.1f ms  %.3e cells/s' % (d['ms_per_step'], d['value']))"
done
for w in 0 6; do
PRX_MAX_WARPS_PER_SM=$w timeout 600 python bench.py --mech drm19 --n 64 --steps 3 --warmup 2 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('drm19 max_warps $w  %.1f ms  %.3e cells/s' % (d['ms_per_step'], d['value']))"
done
