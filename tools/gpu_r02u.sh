#!/bin/bash
# session r02u: wall time of the PMF application (CUDA build, 2 AMR levels) per coarse step with the stock GPU reactors and the B200 ones
mkdir -p gpurun_out
cd /tmp
for integ in ReactorRK64 ReactorRK64B200 ReactorCvode ReactorCvodeB200; do
  extra=""
  [ $integ = ReactorCvode ] && extra="cvode.solve_type=GMRES"
  python /root/repo/tools/pelec_pmf_run.py $integ 20 /tmp/u_$integ.npz --cuda ode.rtol=1e-8 ode.atol=1e-10 cvode.max_order=3 $extra 2>/dev/null | grep PELEC_PMF_DONE | tee -a /root/repo/gpurun_out/r02u_pmf_cuda_walls.txt
done
