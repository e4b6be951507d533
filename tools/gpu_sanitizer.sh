#!/bin/bash
# compute-sanitizer (memcheck) over one small call of every shipped integrator kernel and the fused react_state path; the log goes
# to profiles/ (race / memory checking).  usage (GPU box): tools/gpu_sanitizer.sh <tag>
set -u
TAG=${1:-r02}
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 5 python -m pytest tests/test_gpu_state.py tests/test_gpu_parity2.py -m gpu -q -x \
   -k "react_state_cvode_matches or react_boxes_equals or flat_react or dodecane_cvode_64" > gpurun_out/${TAG}_sanitizer_memcheck.log 2>&1
echo "memcheck rc=$?" | tee -a gpurun_out/${TAG}_sanitizer_memcheck.log
tail -6 gpurun_out/${TAG}_sanitizer_memcheck.log
