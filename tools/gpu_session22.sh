#!/bin/bash
set -u
L=$PWD/pelec_b200/lib/libpelerx_b200
for t in "" _c16 _c12 _c8; do
PRX_LIB_PATH=${L}$t.so timeout 900 python bench.py --mech dodecane_lu --n 32 --steps 2 --warmup 2 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('dodecane lib[$t]  %.1f ms  %.3e cells/s  e2e %.3e' % (d['ms_per_step'], d['value'], d['e2e']['value']))"
done
PRX_LIB_PATH=${L}_c16.so timeout 300 python tools/dbg.py dodecane_lu 2>&1 | tail -1
