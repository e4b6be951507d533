#!/usr/bin/env python
"""Generate tests/golden/react_state_pmf_<mech>_<integrator>.npz from a LIVE, UNMODIFIED PeleC (run in the build container only).

oracle/_ref/libpelec_pmf_<mech>.so (oracle/ref_pelec/Makefile: PeleC + AMReX + PelePhysics + SUNDIALS compiled from
/root/reference) starts the PMF regression case (PC/Exec/RegTests/PMF/pmf-lidryer-cvode.inp, level 0 only, 8 x 8 x 128),
advances it two coarse steps with the application's own driver, then PeleC::react_state (PC/Source/React.cpp:30-362) is
called on the level with S_old = the state before the last step and a non-reacting source built from the application's own
increment.  Inputs and outputs of that call for a subset of cells are the fixture that pins
  * oracle.pyoracle.react_state_oracle (CPU test, tests/test_react_state_golden.py) and
  * prx_react_state_box (GPU test, same file).
One process per integrator (PeleC reads pelec.chem_integrator once): this script re-executes itself.
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
CASE = "/root/reference/Exec/RegTests/PMF"
PARAMS = {"rtol": 1e-8, "atol": 1e-10, "max_order": 3}
DT = 1e-7
NSUB = 640  # cells kept in the fixture


RF = {"omega": 4.0e4, "axis": 2, "axis_loc": (0.11, 0.07, 0.0)}  # rotating-frame variant (pelec.do_rf)


def worker(mech, integ, out):
    rf = out.endswith("_rf.npz")
    lib = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libpelec_pmf_%s.so" % mech))
    args = ["pelec", CASE + "/pmf-lidryer-cvode.inp", "amr.plot_files_output=0", "amr.checkpoint_files_output=0",
            "pelec.xNormPlane.int=100000", "prob.pmf_datafile=" + CASE + "/LiDryer_H2_p1_phi0_4000tu0300.dat",
            "amr.max_level=0", "pelec.v=0", "amr.v=0", "pelec.sum_interval=-1",
            "pelec.chem_integrator=" + integ, "cvode.solve_type=denseAJ_direct", "ode.analytical_jacobian=1",
            "ode.rtol=%r" % PARAMS["rtol"], "ode.atol=%r" % PARAMS["atol"], "cvode.max_order=%d" % PARAMS["max_order"],
            "ode.clean_init_massfrac=0"]
    if rf:
        args += ["pelec.do_rf=1", "pelec.do_mol=1", "pelec.rf_omega=%r" % RF["omega"], "pelec.rf_axis=%d" % RF["axis"],
                 "pelec.rf_axis_x=%r" % RF["axis_loc"][0], "pelec.rf_axis_y=%r" % RF["axis_loc"][1], "pelec.rf_axis_z=%r" % RF["axis_loc"][2]]
    argv = (C.c_char_p * len(args))(*[a.encode() for a in args])
    lib.pelecref_init.argtypes = [C.c_int, C.POINTER(C.c_char_p)]
    lib.pelecref_ncells.restype = C.c_long
    lib.pelecref_coarse_steps.restype = C.c_double
    lib.pelecref_last_dt.restype = C.c_double
    lib.pelecref_get.argtypes = [C.c_int, C.c_int, C.c_void_p]
    lib.pelecref_set.argtypes = [C.c_int, C.c_int, C.c_void_p]
    lib.pelecref_react_state.argtypes = [C.c_int, C.c_double, C.c_int, C.c_void_p]
    lib.pelecref_init(len(args), argv)
    nc, nvar, N = lib.pelecref_ncells(0), lib.pelecref_nvar(), lib.pelecref_num_species()
    idx = (C.c_int * 8)()
    lib.pelecref_indices(idx)
    assert list(idx) == [0, 1, 2, 3, 4, 5, 6, 7] and nvar == 7 + N, (list(idx), nvar)
    lib.pelecref_coarse_steps(2)
    dt_app = lib.pelecref_last_dt(0)

    def get(which, ncomp):
        a = np.zeros((ncomp, nc))
        lib.pelecref_get(0, which, a.ctypes.data)
        return a
    sold, snew_app = get(0, nvar), get(1, nvar)
    rng = np.random.default_rng(3)
    # non-reacting source: the application's own rate of change over its last step (hydro + diffusion + reactions), halved,
    # plus a transverse momentum source so that every term of the kinetic-energy arithmetic is exercised
    aux = 0.5 * (snew_app - sold) / dt_app
    aux[6] = 0.0
    aux[1] += 2.0e5 * sold[0] * rng.uniform(-1, 1, nc)
    aux[2] += 2.0e5 * sold[0] * rng.uniform(-1, 1, nc)
    aux[0] = aux[7:].sum(axis=0)
    res = {}
    for tag, react_init in (("update", 0), ("init", 1)):
        lib.pelecref_set(0, 0, sold.ctypes.data)
        lib.pelecref_set(0, 1, snew_app.ctypes.data)
        lib.pelecref_react_state(0, DT, react_init, None if react_init else aux.ctypes.data)
        res[tag] = (get(1, nvar), get(2, N + 2))
    if rf:
        # every cell of the box with the strongest heat release, with its global indices: rotenrg depends on the position
        lib.pelecref_boxes.argtypes = [C.c_int, C.c_void_p]
        nb = lib.pelecref_nboxes(0)
        bx = np.zeros(6 * nb, dtype=np.int32)
        lib.pelecref_boxes(0, bx.ctypes.data)
        bx = bx.reshape(nb, 6)
        sizes = [int(np.prod(b[3:] - b[:3] + 1)) for b in bx]
        starts = np.concatenate([[0], np.cumsum(sizes)])
        hr = np.abs(res["update"][1][N + 1])
        b = int(np.argmax([hr[starts[i]:starts[i + 1]].max() for i in range(nb)]))
        sel = np.arange(starts[b], starts[b + 1])
        np.savez_compressed(
            out, mech=mech, integrator=integ, dt=DT, atol=PARAMS["atol"], sold=sold[:, sel], aux=aux[:, sel],
            snew_out=res["update"][0][:, sel], I_R=res["update"][1][:, sel], box_lo=bx[b, :3], box_hi=bx[b, 3:],
            omega=RF["omega"], axis=RF["axis"], axis_loc=np.array(RF["axis_loc"]), prob_lo=np.array([0.0, 0.0, 1.0]),
            dx=np.array([0.3125 / 8, 0.3125 / 8, 5.0 / 128]),
            note="PeleC::react_state with pelec.do_rf = 1 (rotating frame), one whole box - tests/golden/make_golden3.py")
        print("%s %s rotating frame: box %d, %d cells, max |heat release| %.3e" % (mech, integ, b, sel.size, hr[sel].max()))
        return
    # subset: whole z-columns (the profile varies along z, cells of a column are 64 apart in the flat order of a box)
    flame = np.argsort(-np.abs(res["update"][1][N + 1]))[: NSUB // 2]  # strongest heat release
    rest = rng.choice(nc, NSUB - flame.size, replace=False)
    sel = np.unique(np.concatenate([flame, rest]))
    np.savez_compressed(
        out, mech=mech, integrator=integ, dt=DT, rtol=PARAMS["rtol"], atol=PARAMS["atol"], max_order=PARAMS["max_order"],
        sold=sold[:, sel], aux=aux[:, sel], snew_in_init=snew_app[:, sel],
        snew_out=res["update"][0][:, sel], I_R=res["update"][1][:, sel],
        snew_out_init=res["init"][0][:, sel], I_R_init=res["init"][1][:, sel],
        note="PeleC::react_state of the unmodified reference (PMF case, level 0) - tests/golden/make_golden3.py")
    hr = res["update"][1][N + 1]
    print("%s %s: %d of %d cells kept, T %.0f..%.0f K, max |heat release| %.3e, dt_app %.3e" % (
        mech, integ, sel.size, nc, sold[6].min(), sold[6].max(), np.abs(hr).max(), dt_app))


def main():
    if len(sys.argv) == 4:
        worker(sys.argv[1], sys.argv[2], sys.argv[3])
        return
    here = os.path.dirname(os.path.abspath(__file__))
    for integ, tag in (("ReactorCvode", "cvode"), ("ReactorRK64", "rk64"), ("ReactorRK64", "rk64_rf")):
        out = os.path.join(here, "react_state_pmf_LiDryer_%s.npz" % tag)
        tmp = "/tmp/pelec_golden_%s" % tag
        os.makedirs(tmp, exist_ok=True)
        p = subprocess.run([sys.executable, os.path.abspath(__file__), "LiDryer", integ, out], cwd=tmp,
                           capture_output=True, text=True)
        print(p.stdout.strip().splitlines()[-1] if p.stdout.strip() else p.stderr[-2000:])


if __name__ == "__main__":
    main()
