"""PeleC::react_state (PC/Source/React.cpp:30-362) pinned against the UNMODIFIED application.

tests/golden/react_state_pmf_LiDryer_{cvode,rk64}.npz hold inputs and outputs of PeleC::react_state called on a live level
of the PMF regression case (oracle/ref_pelec + tests/golden/make_golden3.py; 628 cells through a lean H2 flame, 298-1426 K,
the application's own rate of change as the non-reacting source).  Checked here:

 * CPU (-m "not gpu"): the numpy restatement oracle.pyoracle.react_state_oracle - RK64: the pack / unpack / I_R / heat-release
   arithmetic to rounding (the oracle's RK64 is bit-exact with the reference's); CVODE: the reactor's own bar;
 * GPU (-m gpu): prx_react_state_box (fused load / store phases of the integrator kernels) through the C ABI, same bars.

Nothing here reads /root/reference."""
import os

import numpy as np
import pytest

from oracle.pyoracle import Oracle, react_state_oracle

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def load(tag):
    z = np.load(os.path.join(GOLD, "react_state_pmf_LiDryer_%s.npz" % tag))
    return {k: z[k] for k in z.files}


def bars(g, N):
    """per-cell, per-species reactor bar 10*(atol + rtol*|rhoY|) (tests/util.py) and the scales of the derived quantities"""
    atol, rtol = float(g["atol"]), float(g["rtol"])
    lim = 10.0 * (atol + rtol * np.maximum(np.abs(g["sold"][7:]), np.abs(g["snew_out"][7:])))
    return lim


def compare(g, N, got_snew, got_IR, ref_snew, ref_IR, dt, species_bar, tight):
    """species_bar: [N, ncells] absolute bar on rhoY_k; tight: relative bar on pure pack / unpack arithmetic"""
    assert (np.abs(got_snew[7:] - ref_snew[7:]) <= species_bar).all()
    assert (np.abs(got_snew[6] - ref_snew[6]) / ref_snew[6]).max() <= (1e-6 if species_bar.max() > 1e-14 else 1e-12)
    for comp in range(6):
        sc = np.abs(ref_snew[comp]).max() + 1e-300
        extra = species_bar.sum(axis=0).max() if comp == 0 else 0.0  # rho = sum of the species
        assert np.abs(got_snew[comp] - ref_snew[comp]).max() <= tight * sc + extra, comp
    assert (np.abs(got_IR[:N] - ref_IR[:N]) <= species_bar / dt + tight * np.abs(ref_IR[:N]).max()).all()
    assert np.abs(got_IR[N] - ref_IR[N]).max() <= tight * np.abs(g["sold"][4]).max() / dt
    hscale = np.abs(ref_IR[N + 1]).max()
    h_bar = 3e10 * (species_bar / dt).sum(axis=0)  # |h_k| <= 3e10 erg/g over 300-1500 K for H / O / OH radicals
    assert (np.abs(got_IR[N + 1] - ref_IR[N + 1]) <= h_bar + 1e-9 * hscale).all()


@pytest.mark.parametrize("tag", ["rk64", "cvode"])
def test_restated_react_state_matches_pelec(tag):
    g = load(tag)
    orc = Oracle("LiDryer")
    N, dt = orc.N, float(g["dt"])
    sold, aux = g["sold"], g["aux"]
    snew_in = sold + dt * aux  # React.cpp:89-92 (MultiFab::Copy + Saxpy)
    if tag == "rk64":
        kw = dict(atol=float(g["atol"]))
        species_bar = 1e-13 * np.maximum(np.abs(g["snew_out"][7:]), 1e-6 * g["snew_out"][0][None, :])
        tight = 1e-13
    else:
        kw = dict(rtol=float(g["rtol"]), atol=float(g["atol"]), max_order=int(g["max_order"]), solve_type="denseAJ_direct")
        species_bar = bars(g, N)
        tight = 1e-9
    got_snew, got_IR, _ = react_state_oracle(orc, sold, snew_in, aux, dt, True, tag, **kw)
    compare(g, N, got_snew, got_IR, g["snew_out"], g["I_R"], dt, species_bar, tight)
    # start-up call: S_new untouched, I_R from the state itself with zero source (React.cpp:58-62, 141-142)
    s0 = g["snew_in_init"]
    got_snew, got_IR, _ = react_state_oracle(orc, s0, s0.copy(), np.zeros_like(s0), dt, False, tag, **kw)
    assert np.array_equal(got_snew, g["snew_out_init"])
    gi = dict(g, sold=s0, snew_out=g["snew_out_init"])
    sb = species_bar if tag == "cvode" else 1e-13 * np.maximum(np.abs(s0[7:]), 1e-6 * s0[0][None, :])
    assert (np.abs(got_IR[:N] - g["I_R_init"][:N]) <= sb / dt + tight * np.abs(g["I_R_init"][:N]).max()).all()
    assert np.abs(got_IR[N] - g["I_R_init"][N]).max() <= tight * np.abs(s0[4]).max() / dt
    del gi


def rf_case():
    from oracle.pyoracle import rot_energy
    g = load("rk64_rf")
    lo, hi = g["box_lo"], g["box_hi"]
    shape = tuple(int(hi[d] - lo[d] + 1) for d in range(3))
    c = np.arange(shape[0] * shape[1] * shape[2])
    idx = np.stack([lo[0] + c % shape[0], lo[1] + (c // shape[0]) % shape[1], lo[2] + c // (shape[0] * shape[1])])
    rot = rot_energy(idx, float(g["omega"]), int(g["axis"]), g["axis_loc"], g["prob_lo"], g["dx"])
    return g, shape, rot


def test_restated_react_state_with_rotating_frame_matches_pelec():
    """pelec.do_rf = 1 (React.cpp:123-131, 192-218, 245-300): one whole box of the PMF level with its global indices"""
    g, shape, rot = rf_case()
    assert rot.min() > 0 and rot.max() > 10 * rot.min()  # the term is exercised, and varies over the box
    orc = Oracle("LiDryer")
    dt = float(g["dt"])
    sold, aux = g["sold"], g["aux"]
    got_snew, got_IR, _ = react_state_oracle(orc, sold, sold + dt * aux, aux, dt, True, "rk64", rotenrg=rot, atol=float(g["atol"]))
    assert np.array_equal(got_snew, g["snew_out"]) and np.array_equal(got_IR, g["I_R"])  # bit for bit


@pytest.mark.gpu
def test_gpu_react_state_with_rotating_frame_matches_pelec():
    torch = pytest.importorskip("torch")
    from pelec_b200.reactor import Array4, Box, ReactorB200
    g, shape, rot = rf_case()
    dev = torch.device("cuda:0")
    N, dt = 9, float(g["dt"])
    nx, ny, nz = shape
    lo, hi = [int(v) for v in g["box_lo"]], [int(v) for v in g["box_hi"]]
    r = ReactorB200.create("ReactorRK64", "LiDryer", ode__atol=float(g["atol"]))
    r.init(1, 1)
    r.set_rotating_frame(float(g["omega"]), int(g["axis"]), g["axis_loc"], g["prob_lo"], g["dx"])

    def fab(a):
        return torch.from_numpy(np.ascontiguousarray(a).reshape(a.shape[0], nz, ny, nx).copy()).to(dev)
    sold, aux = g["sold"], g["aux"]
    t_old, t_new, t_src = fab(sold), fab(sold + dt * aux), fab(aux)
    t_IR = torch.zeros((N + 2, nz, ny, nx), dtype=torch.float64, device=dev)
    A = lambda t: Array4(t, lo)  # the FAB starts at the box's global lo
    r.react_state(Box(lo, hi), A(t_old), A(t_new), A(t_src), A(t_IR), dt)
    torch.cuda.synchronize()
    got_snew = t_new.reshape(t_new.shape[0], -1).cpu().numpy()
    got_IR = t_IR.reshape(N + 2, -1).cpu().numpy()
    species_bar = 1e-10 * np.maximum(np.abs(g["snew_out"][7:]), 1e-2 * g["snew_out"][0][None, :])
    compare(g, N, got_snew, got_IR, g["snew_out"], g["I_R"], dt, species_bar, 1e-11)
    # and the term really is in there: without the frame rho*e_int differs (rho*E does not: the term cancels in it when mass
    # is conserved, UEDEN_new = UEDEN_in - KE_in + KE_new + (rho_in - rho_new) rotenrg)
    r.set_rotating_frame(0.0, 2, (0, 0, 0), (0, 0, 0), (1, 1, 1), on=False)
    t_new2 = fab(sold + dt * aux)
    r.react_state(Box(lo, hi), A(t_old), A(t_new2), A(t_src), A(t_IR), dt)
    torch.cuda.synchronize()
    assert (t_new2[5] - t_new[5]).abs().max().item() > 1e-4 * t_new[5].abs().max().item()
    r.close()


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["rk64", "cvode"])
def test_gpu_react_state_matches_pelec(tag):
    torch = pytest.importorskip("torch")
    from pelec_b200.reactor import Array4, Box, ReactorB200
    g = load(tag)
    dev = torch.device("cuda:0")
    N, dt = 9, float(g["dt"])
    sold, aux = g["sold"], g["aux"]
    nc = sold.shape[1]
    snew_in = sold + dt * aux
    if tag == "rk64":
        r = ReactorB200.create("ReactorRK64", "LiDryer", ode__atol=float(g["atol"]))
        # RK64 on the device: the bar of tests/util.py (1e-10 against max(Y_k, 1e-2)), see DESIGN.md 4
        species_bar = 1e-10 * np.maximum(np.abs(g["snew_out"][7:]), 1e-2 * g["snew_out"][0][None, :])
        tight = 1e-11
    else:
        r = ReactorB200.create("ReactorCvode", "LiDryer", ode__rtol=float(g["rtol"]), ode__atol=float(g["atol"]),
                               cvode__max_order=int(g["max_order"]), cvode__solve_type="denseAJ_direct")
        species_bar = bars(g, N)
        tight = 1e-9
    r.init(1, 1)

    def fab(a):
        return torch.from_numpy(np.ascontiguousarray(a).reshape(a.shape[0], 1, 1, nc).copy()).to(dev)
    box = Box((0, 0, 0), (nc - 1, 0, 0))
    t_old, t_new, t_src = fab(sold), fab(snew_in), fab(aux)
    t_IR = torch.zeros((N + 2, 1, 1, nc), dtype=torch.float64, device=dev)
    r.react_state(box, Array4(t_old), Array4(t_new), Array4(t_src), Array4(t_IR), dt)
    torch.cuda.synchronize()
    got_snew = t_new.reshape(t_new.shape[0], -1).cpu().numpy()
    got_IR = t_IR.reshape(N + 2, -1).cpu().numpy()
    compare(g, N, got_snew, got_IR, g["snew_out"], g["I_R"], dt, species_bar, tight)
    # start-up call
    s0 = g["snew_in_init"]
    t_new = fab(s0)
    t_src = torch.zeros_like(t_new)
    t_IR = torch.zeros((N + 2, 1, 1, nc), dtype=torch.float64, device=dev)
    r.react_state(box, Array4(t_new), Array4(t_new), Array4(t_src), Array4(t_IR), dt, do_update=False)
    torch.cuda.synchronize()
    assert np.array_equal(t_new.reshape(t_new.shape[0], -1).cpu().numpy(), g["snew_out_init"])
    got_IR = t_IR.reshape(N + 2, -1).cpu().numpy()
    sb = species_bar if tag == "cvode" else 1e-10 * np.maximum(np.abs(s0[7:]), 1e-2 * s0[0][None, :])
    assert (np.abs(got_IR[:N] - g["I_R_init"][:N]) <= sb / dt + tight * np.abs(g["I_R_init"][:N]).max()).all()
    assert np.abs(got_IR[N] - g["I_R_init"][N]).max() <= tight * np.abs(s0[4]).max() / dt
    r.close()
