#!/bin/bash
# session r02ao: compute-sanitizer memcheck over the code added in the last sessions: ReactorArkode (all dense-output degrees, 1-D
# react), the multi-stream slab pipeline with masked cells (both reactors)
mkdir -p gpurun_out
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 5 python -m pytest tests/test_arkode.py tests/test_gpu_parity.py -m gpu -q -x \
   -k "flat_react or unknown_method or (other_butcher and (42 or 52 or 60)) or slab_pipeline" > gpurun_out/r02ao_sanitizer_memcheck.log 2>&1
echo "memcheck rc=$?" | tee -a gpurun_out/r02ao_sanitizer_memcheck.log
tail -6 gpurun_out/r02ao_sanitizer_memcheck.log
