#!/usr/bin/env python3
"""Round-2 fixtures, generated with the UNMODIFIED reference (oracle/_ref/libpeleref_<mech>.so):

  ignition_LiDryer.npz      the reference's own zero-D known answer (PC/Exec/RegTests/zeroD: H2/O2/N2, Y = 0.06/0.5/0.44,
                            1000 K, 1 atm, constant volume, LiDryer): 1000 successive ReactorCvode::react calls of 1e-6 s
                            (state after every call) and 1000 successive ReactorRK64::react calls, plus the columns of
                            PC/Exec/RegTests/zeroD/cantera_soln (every 1e-5 s).
  reactor2_dodecane_lu.npz  64 dodecane_lu cells (32 ReactEval states + 32 flame-table states of
                            pmf_states_dodecane_lu.npz): ReactorCvode (ReactEval and PeleC tolerances) and - on the flame
                            states, where the reference's RK64 stays finite - ReactorRK64.
  flat_<mech>.npz           the reference's 1-D react(Real*, ..., ncells = 1) (ReactorCvode.cpp:1491-1689,
                            ReactorRK64.cpp:22-164) applied cell by cell: output vectors and the returned int.

Run in the build container:  python tests/golden/make_golden2.py [ignition] [dodecane] [flat]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.pyoracle import BoxState, Reference, reacteval_states  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def ignition():
    ref = Reference("LiDryer")
    Y = np.zeros((1, ref.N))
    Y[0, ref.species.index("H2")], Y[0, ref.species.index("O2")], Y[0, ref.species.index("N2")] = 0.06, 0.5, 0.44
    rho, e = ref.pyt2re(np.array([1013250.0]), Y, np.array([1000.0]))
    st0 = BoxState((Y * rho[:, None]).T.copy(), np.array([1000.0]), rho * e)
    ncall, dt = 1000, 1e-6
    out = {"rY0": st0.rY, "T0": st0.T, "rE0": st0.rE, "dt": dt, "ncall": ncall}
    for name in ("cvode", "rk64"):
        s = st0.copy()
        T, rY, FC = np.zeros(ncall), np.zeros((ncall, ref.N)), np.zeros(ncall)
        for n in range(ncall):
            if name == "cvode":
                ref.react_cvode(s, dt, time=n * dt, rtol=1e-10, atol=1e-10, max_order=4, solve_type="denseAJ_direct")
            else:
                ref.react_rk64(s, dt, time=n * dt, atol=1e-10)
            T[n], rY[n], FC[n] = s.T[0], s.rY[:, 0], s.FC[0]
        out.update({name + "_T": T, name + "_rY": rY, name + "_FC": FC})
        print(name, "T(0.13 ms) %.6f T(0.14 ms) %.6f T(1 ms) %.6f  work %d" % (T[129], T[139], T[-1], FC.sum()))
    can = np.loadtxt("/root/reference/Exec/RegTests/zeroD/cantera_soln", skiprows=1)
    out.update(cantera_t=can[:, 0], cantera_T=can[:, 1], cantera_Y_H2=can[:, 2], cantera_Y_O2=can[:, 3], cantera_Y_OH=can[:, 4])
    path = os.path.join(OUT, "ignition_LiDryer.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes")


def dodecane():
    mech = "dodecane_lu"
    ref = Reference(mech)
    a = reacteval_states(ref, (1, 32, 1))
    z = np.load(os.path.join(OUT, "pmf_states_%s.npz" % mech))
    idx = np.argsort(z["x"])[:: len(z["x"]) // 32][:32]  # 32 positions across the flame, cold reactants to hot products
    st0 = BoxState(np.concatenate([a.rY, z["rY"][:, idx]], axis=1), np.concatenate([a.T, z["T"][idx]]),
                   np.concatenate([a.rE, z["rE"][idx]]))
    out = {"rY0": st0.rY, "T0": st0.T, "rE0": st0.rE, "dt": 1e-6}
    for i, kw in enumerate((dict(rtol=1e-10, atol=1e-10, max_order=4), dict(rtol=1e-6, atol=1e-10, max_order=2))):
        s = st0.copy()
        ref.react_cvode(s, 1e-6, solve_type="denseAJ_direct", **kw)
        out.update({"cv%d_rY" % i: s.rY, "cv%d_T" % i: s.T, "cv%d_rE" % i: s.rE, "cv%d_FC" % i: s.FC})
        print("cvode", kw, "FC", s.FC.min(), s.FC.max())
    # RK64: flame states only, dt = 1e-7 (the ReactEval states drive the reference's own RK64 to NaN)
    f = st0.take(np.arange(32, 64))
    s = f.copy()
    ref.react_rk64(s, 1e-7, atol=1e-10)
    ok = np.isfinite(s.rY).all(axis=0) & np.isfinite(s.T)
    print("rk64 finite cells", ok.sum(), "of", len(ok), "steps", s.FC.min(), s.FC.max())
    out.update(rk64_dt=1e-7, rk64_idx=np.arange(32, 64)[ok], rk64_rY=s.rY[:, ok], rk64_T=s.T[ok], rk64_rE=s.rE[ok], rk64_FC=s.FC[ok])
    path = os.path.join(OUT, "reactor2_%s.npz" % mech)
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes")


def flat():
    for mech, ny in (("LiDryer", 16), ("drm19", 8)):
        ref = Reference(mech)
        st = reacteval_states(ref, (1, ny, 1))
        rng = np.random.default_rng(777)
        N = ref.N
        y0 = np.concatenate([st.rY.T, (st.T + rng.uniform(-20, 20, st.ncells))[:, None]], axis=1).copy()  # [ncells, N+1]
        ysrc = (0.01 * rng.uniform(-1, 1, (st.ncells, N)) * st.rY.T / 1e-6).copy()
        e0 = st.rE.copy()
        esrc = 0.005 * rng.uniform(-1, 1, st.ncells) * np.abs(st.rE) / 1e-6
        out = {"y0": y0, "ysrc": ysrc, "e0": e0, "esrc": esrc, "dt": 1e-6}
        # parameters are global ParmParse state: set them through one Array4 call on a scratch state
        for key, setter in (("ReactorCvode", lambda: ref.react_cvode(st.take([0]), 1e-9, rtol=1e-10, atol=1e-10, max_order=4)),
                            ("ReactorRK64", lambda: ref.react_rk64(st.take([0]), 1e-9, atol=1e-10))):
            setter()
            y, e, ret = y0.copy(), e0.copy(), np.zeros(st.ncells, dtype=np.int64)
            for c in range(st.ncells):  # ncells = 1 per call: per-cell semantics (see include/pelerx_b200.h)
                yc, ec = y[c:c + 1].copy(), e[c:c + 1].copy()
                ret[c] = ref.react_flat(key, yc, ysrc[c:c + 1].copy(), ec, esrc[c:c + 1].copy(), 1e-6, 0.0)
                y[c], e[c] = yc[0], ec[0]
            out.update({key + "_y": y, key + "_e": e, key + "_ret": ret})
            print(mech, key, "ret", ret)
        path = os.path.join(OUT, "flat_%s.npz" % mech)
        np.savez_compressed(path, **out)
        print(path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    what = sys.argv[1:] or ["ignition", "dodecane", "flat"]
    if "ignition" in what:
        ignition()
    if "flat" in what:
        flat()
    if "dodecane" in what:
        dodecane()
