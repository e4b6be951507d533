#!/usr/bin/env python
"""TEST INFRASTRUCTURE: run the UNMODIFIED PeleC application (PMF regression case, oracle/_ref/libpelec_pmf_<mech>.so built by
oracle/ref_pelec/Makefile) for n coarse time steps with a given pelec.chem_integrator and dump every level's state.

  python tools/pelec_pmf_run.py <integrator> <nsteps> <out.npz> [--cuda] [name=value ...]

The library also holds integration/ReactorB200.cpp, so <integrator> may be ReactorCvodeB200 / ReactorRK64B200: the reference's
own factory then creates this repo's reactor and PeleC::react_state (PC/Source/React.cpp:222-230) drives the CUDA kernels
through the reference's virtual react() - the drop-in, in the application.  Used by tests/test_pelec_dropin.py (two
processes: stock vs B200).  Reads its case data from oracle/_ref/pmf_case/ (copied there by the Makefile), never /root/reference.
"""
import ctypes as C
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    integ, nsteps, out = sys.argv[1], int(sys.argv[2]), os.path.abspath(sys.argv[3])
    extra = sys.argv[4:]
    mech = "LiDryer"
    case = os.path.join(ROOT, "oracle", "_ref", "pmf_case")
    # --cuda: the CUDA build of the application (oracle/ref_pelec/Makefile.cuda; AMReX, PeleC, PelePhysics compiled with nvcc
    # for sm_100a): state in device memory, the shim's AMREX_USE_GPU branch hands device pointers to prx_react_box
    cuda = "--cuda" in extra
    if cuda:
        extra = [e for e in extra if e != "--cuda"] + ["amrex.the_arena_init_size=0", "amrex.abort_on_out_of_gpu_memory=1"]
    lib = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libpelec_pmf_%s%s.so" % ("cuda_" if cuda else "", mech)))
    os.chdir(tempfile.mkdtemp(prefix="pelec_pmf_"))
    args = ["pelec", os.path.join(case, "pmf-lidryer-cvode.inp"), "amr.plot_files_output=0", "amr.checkpoint_files_output=0",
            "pelec.xNormPlane.int=100000", "prob.pmf_datafile=" + os.path.join(case, "LiDryer_H2_p1_phi0_4000tu0300.dat"),
            "pelec.v=0", "amr.v=0", "pelec.sum_interval=-1", "pelec.chem_integrator=" + integ,
            "ode.clean_init_massfrac=0"]
    if not any(e.startswith("cvode.solve_type=") for e in extra):  # default of these runs; "cvode.solve_type=GMRES" = the shipped input
        args += ["cvode.solve_type=denseAJ_direct", "ode.analytical_jacobian=1"]
    args += extra
    argv = (C.c_char_p * len(args))(*[a.encode() for a in args])
    lib.pelecref_init.argtypes = [C.c_int, C.POINTER(C.c_char_p)]
    lib.pelecref_ncells.restype = C.c_long
    lib.pelecref_coarse_steps.restype = C.c_double
    lib.pelecref_get.argtypes = [C.c_int, C.c_int, C.c_void_p]
    lib.pelecref_boxes.argtypes = [C.c_int, C.c_void_p]
    finest = lib.pelecref_init(len(args), argv)
    t0 = time.perf_counter()
    t = lib.pelecref_coarse_steps(nsteps)
    wall = time.perf_counter() - t0
    res = {"time": t, "wall_s": wall, "finest": finest}
    nvar = lib.pelecref_nvar()
    for lev in range(finest + 1):
        nc = lib.pelecref_ncells(lev)
        s = np.zeros((nvar, nc))
        lib.pelecref_get(lev, 1, s.ctypes.data)
        nb = lib.pelecref_nboxes(lev)
        bx = np.zeros(6 * nb, dtype=np.int32)
        lib.pelecref_boxes(lev, bx.ctypes.data)
        ir = np.zeros((lib.pelecref_ncomp(lev, 2), nc))
        lib.pelecref_get(lev, 2, ir.ctypes.data)
        res["S%d" % lev], res["boxes%d" % lev], res["IR%d" % lev] = s, bx, ir
    np.savez(out, **res)
    sys.stdout.flush()
    print("PELEC_PMF_DONE %s steps=%d time=%.6e wall=%.2fs levels=%d" % (integ, nsteps, t, wall, finest + 1))


if __name__ == "__main__":
    main()
