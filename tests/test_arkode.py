"""ReactorArkode (PP/Source/Reactions/ReactorArkode.cpp:135-260, the ERKStep default) behind the C ABI: the box-coupled adaptive
explicit Runge-Kutta integration of ARKODE - one step size and one WRMS error norm for the whole box, PID / PI controller whose
history carries over from one react call to the next, Hermite interpolation back to tout - restated for the GPU
(pelec_b200/csrc/prx_arkode.cuh).  Fixtures: three successive calls of one reference reactor object on three different boxes
(tests/golden/make_golden5.py, unmodified reference).

Bar: the step-size sequence depends on box-wide norms that are summed in a different order on the device, so results agree to
rounding-amplified-by-the-controller rather than bit for bit: |d rhoY_k| <= 1e-5 * (atol + rtol |rhoY_k|) (measured: 1e-7 of the
integrator's own tolerance), |dT| / T <= 1e-11, and the SAME number of right-hand-side evaluations (FC, react's return value) -
i.e. the same accepted / rejected step sequence."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
CASES = {"default": dict(rtol=1e-6, atol=1e-10, rk_method=40, rk_controller=0, typ=False),
         "bs_pi_typ": dict(rtol=1e-8, atol=1e-10, rk_method=30, rk_controller=1, typ=True),
         "arkstep": dict(rtol=1e-6, atol=1e-10, rk_method=40, rk_controller=0, typ=False, use_erkstep=0)}


@pytest.mark.parametrize("mech", ["LiDryer", "drm19"])
@pytest.mark.parametrize("tag", sorted(CASES))
def test_arkode_matches_reference_over_successive_calls(mech, tag):
    from oracle.pyoracle import BoxState
    from pelec_b200.reactor import ReactorB200
    z = np.load(os.path.join(GOLD, "arkode_%s.npz" % mech))
    kw = CASES[tag]
    dt = float(z["dt"])
    r = ReactorB200.create("ReactorArkode", mech, ode__rtol=kw["rtol"], ode__atol=kw["atol"], ode__rk_method=kw["rk_method"],
                           ode__rk_controller=kw["rk_controller"], ode__use_erkstep=kw.get("use_erkstep", 1))
    r.init(1, 1)
    N = z[tag + "_0_rY0"].shape[0]
    if kw["typ"]:
        r.set_typ_vals_ode(np.concatenate([np.full(N, 1e-4), [2000.0]]))
    worst = []
    for i in range(3):
        pre = "%s_%d_" % (tag, i)
        s = BoxState(z[pre + "rY0"].copy(), z[pre + "T0"].copy(), z[pre + "rE0"].copy(), z[pre + "rYsrc"].copy(), z[pre + "rEsrc"].copy())
        ret = r.react_host(s.ncells, 1, 1, s.rY, s.rYsrc, s.T, s.rE, s.rEsrc, s.FC, s.mask, dt, 0.0)
        nfe = int(z[tag + "_nfe"][i])
        lim = kw["atol"] + kw["rtol"] * np.abs(z[pre + "rY"])
        dy = float((np.abs(s.rY - z[pre + "rY"]) / lim).max())
        dT = float((np.abs(s.T - z[pre + "T"]) / z[pre + "T"]).max())
        worst.append((dy, dT, int(s.FC[0]), nfe))
        assert np.all(s.FC == s.FC[0]) and np.array_equal(z[pre + "FC"], np.full(s.ncells, float(nfe)))
        assert int(s.FC[0]) == nfe, worst          # same accepted / rejected step sequence
        assert r.last_react_return() == nfe
        assert dy <= 1e-5 and dT <= 1e-11, worst
        assert np.array_equal(s.rE, z[pre + "rE"])  # rho*e advances by dt * source exactly (box_unflatten)
        del ret
    print("arkode %s %s: (|d rhoY| / tol, |dT|/T, FC, nfe) per call: %s" % (mech, tag, worst))
    r.close()


def test_arkode_unsupported_options_fail_loudly():
    from pelec_b200.reactor import PrxError, ReactorB200
    r = ReactorB200.create("ReactorArkode", "LiDryer", ode__rk_controller=3)  # explicit Gustafsson
    with pytest.raises(PrxError):
        r.init(1, 1)


def test_arkode_unknown_method_is_the_default_table():
    """ReactorArkode.cpp:86-89: an ode.rk_method outside the list selects Zonneveld 5-3-4"""
    from oracle.pyoracle import BoxState
    from pelec_b200.reactor import ReactorB200
    z = np.load(os.path.join(GOLD, "arkode_LiDryer.npz"))
    out = []
    for method in (40, 99):
        r = ReactorB200.create("ReactorArkode", "LiDryer", ode__rtol=1e-6, ode__atol=1e-10, ode__rk_method=method)
        r.init(1, 1)
        pre = "default_0_"
        s = BoxState(z[pre + "rY0"].copy(), z[pre + "T0"].copy(), z[pre + "rE0"].copy(), z[pre + "rYsrc"].copy(), z[pre + "rEsrc"].copy())
        r.react_host(s.ncells, 1, 1, s.rY, s.rYsrc, s.T, s.rE, s.rEsrc, s.FC, s.mask, float(z["dt"]), 0.0)
        out.append((s.rY.copy(), s.T.copy(), r.last_react_return()))
        r.close()
    assert np.array_equal(out[0][0], out[1][0]) and np.array_equal(out[0][1], out[1][1]) and out[0][2] == out[1][2]


@pytest.mark.parametrize("mech", ["LiDryer", "drm19"])
def test_arkode_flat_react_matches_array4_fixture(mech):
    """ReactorArkode's 1-D react (ReactorArkode.cpp:266-380) runs the same box-coupled integration on cell-major vectors whose
    temperature component is taken as given.  The fixture's states carry T consistent with (rho e, rho Y), so the first call of
    the Array4 fixture is the known answer: same number of right-hand-side evaluations (react's return value), rho Y and T to
    the Array4 test's bar (the Array4 path re-solves T from the energy at both ends, the 1-D path returns the integrated T:
    compared to 1e-9), rho e advanced by its source."""
    from pelec_b200.reactor import ReactorB200
    z = np.load(os.path.join(GOLD, "arkode_%s.npz" % mech))
    kw = CASES["default"]
    dt = float(z["dt"])
    r = ReactorB200.create("ReactorArkode", mech, ode__rtol=kw["rtol"], ode__atol=kw["atol"], ode__rk_method=kw["rk_method"],
                           ode__rk_controller=kw["rk_controller"])
    r.init(1, 1)
    pre = "default_0_"
    rY0, T0, rE0, rYsrc, rEsrc = (z[pre + k] for k in ("rY0", "T0", "rE0", "rYsrc", "rEsrc"))
    N, nc = rY0.shape
    y = np.ascontiguousarray(np.concatenate([rY0, T0[None, :]], axis=0).T)  # [cell][N + 1]
    ysrc = np.ascontiguousarray(rYsrc.T)
    e, esrc, fc = rE0.copy(), rEsrc.copy(), np.zeros(nc)
    r.react_flat_host(y, ysrc, e, esrc, dt, 0.0, nc, fc=fc)
    nfe = int(z["default_nfe"][0])
    assert r.last_react_return() == nfe and np.all(fc == nfe)
    lim = kw["atol"] + kw["rtol"] * np.abs(z[pre + "rY"])
    dy = float((np.abs(y[:, :N].T - z[pre + "rY"]) / lim).max())
    dT = float((np.abs(y[:, N] - z[pre + "T"]) / z[pre + "T"]).max())
    print("arkode flat %s: |d rhoY| / tol %.2e, |dT|/T %.2e, nfe %d" % (mech, dy, dT, nfe))
    assert dy <= 1e-3 and dT <= 1e-9
    assert np.array_equal(e, rE0 + dt * rEsrc)
    r.close()


@pytest.mark.parametrize("method", [31, 41, 42, 43, 50, 51, 52, 53, 54, 60, 80])
def test_arkode_other_butcher_tables_match_reference(method):
    """Every other explicit table ode.rk_method selects (ReactorArkode.cpp:25-90; tables generated from SUNDIALS' definitions by
    tools/extract_erk_tables.py).  Order-5 tables interpolate back to tout with ARKODE's quartic Hermite polynomial, which costs
    one more right-hand side (arkode_interp.c:578-600), the order-6 / 8 tables (60: Verner 8-5-6, 80: Fehlberg 13-7-8) with the
    quintic one, which costs four (:648-690) - part of the count compared here.  Table 43's
    first and third call run into stages outside the physical range (NaN error norm): the reference's controller treats the
    NaN as a tiny error and the error test as failed; the same sequence of RHS evaluations is the test.
    Fixture: tests/golden/make_golden6.py (unmodified reference, three successive calls of one reactor object)."""
    from oracle.pyoracle import BoxState
    from pelec_b200.reactor import ReactorB200
    z = np.load(os.path.join(GOLD, "arkode_tables_LiDryer.npz"))
    dt = float(z["dt"])
    rtol, atol = 1e-6, 1e-10
    r = ReactorB200.create("ReactorArkode", "LiDryer", ode__rtol=rtol, ode__atol=atol, ode__rk_method=method, ode__rk_controller=0)
    r.init(1, 1)
    worst = []
    for i in range(3):
        pi, po = "in_%d_" % i, "m%d_%d_" % (method, i)
        s = BoxState(z[pi + "rY0"].copy(), z[pi + "T0"].copy(), z[pi + "rE0"].copy(), z[pi + "rYsrc"].copy(), z[pi + "rEsrc"].copy())
        r.react_host(s.ncells, 1, 1, s.rY, s.rYsrc, s.T, s.rE, s.rEsrc, s.FC, s.mask, dt, 0.0)
        nfe = int(z["m%d_nfe" % method][i])
        lim = atol + rtol * np.abs(z[po + "rY"])
        dy = float((np.abs(s.rY - z[po + "rY"]) / lim).max())
        dT = float((np.abs(s.T - z[po + "T"]) / z[po + "T"]).max())
        worst.append((dy, dT, int(s.FC[0]), nfe))
        assert int(s.FC[0]) == nfe and r.last_react_return() == nfe, worst
        assert dy <= 1e-5 and dT <= 1e-11, worst
        assert np.array_equal(s.rE, z[po + "rE"])
    print("arkode rk_method %d: (|d rhoY| / tol, |dT|/T, FC, nfe) per call: %s" % (method, worst))
    r.close()
