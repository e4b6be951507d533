"""the shim's AMREX_USE_GPU branch - device pointers and the caller's stream, the mode PeleC's GPU builds use
(integration/ReactorB200.cpp `#ifdef AMREX_USE_GPU` -> prx_react_box) - compiled with nvcc against the CUDA build of the
UNMODIFIED reference (oracle/ref_gpu: AMReX + PelePhysics + SUNDIALS for sm_100a), created by the reference's own factory as
ReactorCvodeB200 / ReactorRK64B200 and driven through the reference's virtual react(box, Array4..., gpuStream_t) on
device-resident FABs.  Checked against the pinned CPU oracle; the reference's own GPU ReactorRK64 runs beside it.

Needs oracle/_ref/libpeleref_gpu_LiDryer.so (built in the build container; travels to the GPU box)."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "oracle", "_ref", "libpeleref_gpu_LiDryer.so")


def run(reactor, n=16, extra=()):
    p = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ref_gpu_bench.py"), "--mech", "LiDryer", "--n", str(n),
                        "--reactor", reactor, "--reps", "1", "--check"] + list(extra), capture_output=True, text=True, timeout=600)
    assert p.stdout.strip(), p.stderr[-3000:]
    return json.loads(p.stdout.strip().splitlines()[-1])


@pytest.mark.skipif(not os.path.exists(LIB), reason="oracle/_ref/libpeleref_gpu_LiDryer.so not built")
def test_b200_reactors_in_the_cuda_build_of_the_reference():
    d = run("ReactorCvodeB200")
    assert d["impl"] == "b200-in-reference-gpu-build" and d["react_return"] == 20  # ReactorCvode.cpp:1463-1464
    assert d["check"]["max_drY_over_bar"] <= 1.0 and d["check"]["max_rel_dT_vs_cpu_oracle"] <= 1e-6, d["check"]
    assert 20 < d["mean_FC"] < 80  # per-cell nfe + nfeLS like the CPU path, not the box-wide count of the GPU reference
    d = run("ReactorRK64B200")
    assert d["check"]["max_drY_over_rho_vs_cpu_oracle"] <= 1e-12 and d["check"]["max_rel_dT_vs_cpu_oracle"] <= 1e-10, d["check"]
    # the reference's own GPU RK64 through the same door: bit-level agreement with the CPU oracle
    d = run("ReactorRK64")
    assert d["check"]["max_drY_over_rho_vs_cpu_oracle"] <= 1e-13, d["check"]
