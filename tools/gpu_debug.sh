#!/bin/bash
set -u
mkdir -p gpurun_out
which gdb strace ltrace
for v in 0 1; do
echo "== variant $v"
PRX_BDF_VARIANT=$v CUDA_LAUNCH_BLOCKING=1 timeout 300 python -X faulthandler tools/dbg.py 2>&1 | tail -20
echo "rc=${PIPESTATUS[0]}"
done
PRX_BDF_VARIANT=1 strace -f -o gpurun_out/strace.txt -e trace=exit_group,write python tools/dbg.py > /dev/null 2>&1
tail -5 gpurun_out/strace.txt
PRX_BDF_VARIANT=1 gdb -batch -ex "break exit" -ex "break _exit" -ex run -ex bt --args python tools/dbg.py 2>&1 | tail -40
