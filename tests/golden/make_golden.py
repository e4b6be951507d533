#!/usr/bin/env python3
"""Generate tests/golden/*.npz with the UNMODIFIED reference (oracle/_ref/libpeleref_<mech>.so).

Run in the build container (needs /root/reference to have been compiled by oracle/ref/Makefile):
    python tests/golden/make_golden.py
Each fixture holds the seven boundary arrays of ReactorBase::react before and after one call,
for the reference's ReactorRK64 and ReactorCvode, on ReactEval's synthetic states
(PP/Testing/Exec/ReactEval/GPU_misc.H:13-82) plus seeded external sources, and unit-level
vectors (wdot, Jacobian, T-Newton, fKernelSpec).  Small on purpose (a few hundred KB in total).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.pyoracle import Reference, reacteval_states  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
CASES = {
    # mech: (ny, [(reactor, kwargs)...])
    "LiDryer": 32,
    "drm19": 16,
    "dodecane_lu": 4,
}
CVODE_SETTINGS = [
    dict(rtol=1e-10, atol=1e-10, max_order=4, solve_type="denseAJ_direct"),  # ReactEval inputs.3d_Array4 tolerances
    dict(rtol=1e-6, atol=1e-10, max_order=2, solve_type="denseAJ_direct"),   # PeleC defaults (ReactorCvode.H:179-191)
    dict(rtol=1e-8, atol=1e-12, max_order=5, solve_type="dense_direct"),
]


def main():
    for mech, ny in CASES.items():
        ref = Reference(mech)
        st0 = reacteval_states(ref, (1, ny, 1))
        rng = np.random.default_rng(12345)
        # external sources that exercise rYsrc / rEsrc but stay physical: at most +-2 % of each species'
        # own density over dt (plus a trace source for absent species), +-1 % of rho*e over dt
        dt = 1e-6
        rho = st0.rY.sum(axis=0)
        st0.rYsrc = 0.02 * rng.uniform(-1, 1, size=st0.rY.shape) * st0.rY / dt \
            + 1e-8 * rng.uniform(0, 1, size=st0.rY.shape) * rho[None, :] / dt
        st0.rEsrc = 0.01 * rng.uniform(-1, 1, size=st0.ncells) * np.abs(st0.rE) / dt
        Y = (st0.rY / rho).T.copy()
        out = {"rY0": st0.rY, "T0": st0.T, "rE0": st0.rE, "rYsrc": st0.rYsrc, "rEsrc": st0.rEsrc, "dt": 1e-6}
        # unit-level vectors
        n = 16
        T = rng.uniform(300, 3000, n)
        Yr = rng.random((n, ref.N)) ** 3
        Yr /= Yr.sum(1, keepdims=True)
        rh, e = ref.pyt2re(np.full(n, 1013250.0), Yr, T)
        out.update(u_T=T, u_Y=Yr, u_rho=rh, u_e=e, u_wdot=ref.rty2wdot(rh, T, Yr), u_jac=ref.rty2jac(rh, T, Yr),
                   u_Tsolve=ref.ey2t(e, Yr, T + rng.uniform(-200, 200, n)))
        rng2 = np.random.default_rng(999)
        out["u_Tguess"] = T + np.random.default_rng(12345 + 1).uniform(-200, 200, n)
        out["u_Tsolve"] = ref.ey2t(e, Yr, out["u_Tguess"])
        ei, cv, em, hi = ref.thermo(T, Yr)
        out.update(u_ei=ei, u_cv=cv, u_emix=em, u_hi=hi)
        yv = np.concatenate([Yr * rh[:, None], out["u_Tguess"][:, None]], axis=1)
        usrc = rng2.normal(size=(n, ref.N)) * 1e-3
        uesrc = rng2.normal(size=n) * 1e3
        out.update(u_y=yv, u_src=usrc, u_esrc=uesrc, u_ydot=ref.fkernelspec(1e-7, yv, rh * e, uesrc, usrc))
        # reactors
        if mech != "dodecane_lu":  # the reference's RK64 itself blows up (NaN) on these dodecane states
            s = st0.copy()
            ref.react_rk64(s, 1e-6, atol=1e-10)
            out.update(rk64_rY=s.rY, rk64_T=s.T, rk64_rE=s.rE, rk64_FC=s.FC)
        for i, kw in enumerate(CVODE_SETTINGS):
            s = st0.copy()
            ref.react_cvode(s, 1e-6, **kw)
            out.update({"cv%d_rY" % i: s.rY, "cv%d_T" % i: s.T, "cv%d_rE" % i: s.rE, "cv%d_FC" % i: s.FC})
        # mask == -1 behaviour of the CVODE reactor
        s = st0.copy()
        s.mask[::3] = -1
        ref.react_cvode(s, 1e-6, **CVODE_SETTINGS[1])
        out.update(cvmask_mask=s.mask, cvmask_rY=s.rY, cvmask_T=s.T, cvmask_rE=s.rE, cvmask_FC=s.FC)
        path = os.path.join(OUT, "reactor_%s.npz" % mech)
        np.savez_compressed(path, **out)
        print(path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
