#!/bin/bash
# session r02ak: RK64 mask fix + slab-pipeline test, the other ERK tables of ReactorArkode, 1-D Arkode react
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_arkode.py -m gpu -q -s -k "slab_pipeline or arkode" 2>&1 | tail -24 | tee gpurun_out/r02ak_pytest_new.log
