#!/bin/bash
# session r02ab: compute-sanitizer memcheck over the kernels added in this session
mkdir -p gpurun_out
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 5 python -m pytest tests/test_eos_transport.py tests/test_react_state_golden.py tests/test_extra_mechanisms.py -m gpu -q -x \
   -k "padded or rotating or fixture or BurkeDryer" > gpurun_out/r02ab_sanitizer_memcheck.log 2>&1
echo "memcheck rc=$?" | tee -a gpurun_out/r02ab_sanitizer_memcheck.log
tail -8 gpurun_out/r02ab_sanitizer_memcheck.log
