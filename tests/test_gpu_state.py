"""GPU tests (run with -m gpu) of the entry points either side of the reactor call:

 * prx_react_state_box / prx_react_state_boxes: the body of PeleC::react_state (PC/Source/React.cpp:140-343) fused around the
   integrator kernel - frcEExt in the load phase; S_new, I_R and the heat release in the store phase - against a numpy
   restatement of React.cpp (oracle.pyoracle.react_state_oracle) whose reactor is the pinned oracle;
 * the same body through the unfused pack / unpack kernels the cooperative integrator uses;
 * prx_react_boxes: several boxes in one call (the caller's MFIter loop) must give exactly what box-by-box calls give.

Nothing here reads /root/reference."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle.pyoracle import Oracle, react_state_oracle, reacteval_states  # noqa: E402
from pelec_b200.reactor import Array4, Box, ReactorB200  # noqa: E402

DT = 1e-6
CV = {"ode__rtol": 1e-8, "ode__atol": 1e-10, "cvode__max_order": 3, "cvode__solve_type": "denseAJ_direct"}
CVO = dict(rtol=1e-8, atol=1e-10, max_order=3, solve_type="denseAJ_direct")


def make_state(orc, shape, seed):
    """S_old, non-reacting sources and S_new = S_old + dt*sources for an (nx, ny, nz) box: [NVAR, ncells] arrays."""
    nx, ny, nz = shape
    st = reacteval_states(orc, shape)
    rng = np.random.default_rng(seed)
    N, nc = orc.N, st.ncells
    nvar = 7 + N
    rho = st.rY.sum(axis=0)
    vel = rng.normal(size=(3, nc)) * 3.0e3  # cm/s
    sold = np.zeros((nvar, nc))
    sold[0] = rho
    sold[1:4] = rho * vel
    sold[5] = st.rE
    sold[4] = st.rE + 0.5 * rho * (vel ** 2).sum(axis=0)
    sold[6] = st.T + rng.uniform(-5, 5, nc)
    sold[7:] = st.rY
    nonrs = np.zeros((nvar, nc))
    nonrs[7:] = 0.01 * rng.uniform(-1, 1, (N, nc)) * st.rY / DT + 1e-9 * rng.uniform(0, 1, (N, nc)) * rho / DT
    nonrs[0] = nonrs[7:].sum(axis=0)
    nonrs[1:4] = 0.01 * rng.uniform(-1, 1, (3, nc)) * np.abs(sold[1:4]).max() / DT
    nonrs[4] = 0.005 * rng.uniform(-1, 1, nc) * sold[4] / DT
    nonrs[5] = nonrs[4] * 0.9
    snew = sold + DT * nonrs
    return sold, snew, nonrs


def to_fab(a, shape, dev, pad=(0, 0, 0), fill=0.0):
    """[ncomp, ncells] -> tensor [ncomp, nz+2pz, ny+2py, nx+2px] holding the box in its interior"""
    nx, ny, nz = shape
    px, py, pz = pad
    t = torch.full((a.shape[0], nz + 2 * pz, ny + 2 * py, nx + 2 * px), fill, dtype=torch.float64, device=dev)
    t[:, pz:pz + nz, py:py + ny, px:px + nx] = torch.from_numpy(a.reshape(a.shape[0], nz, ny, nx).copy()).to(dev)
    return t


def from_fab(t, shape, pad=(0, 0, 0)):
    nx, ny, nz = shape
    px, py, pz = pad
    return t[:, pz:pz + nz, py:py + ny, px:px + nx].reshape(t.shape[0], -1).cpu().numpy()


def check_against_oracle(orc, sold, got_snew, got_IR, ref_snew, ref_IR, do_update):
    N = orc.N
    atol, rtol = CVO["atol"], CVO["rtol"]
    # the reactor's own parity bar on the species it returns: 10 * (atol + rtol*|rhoY|); T to 1e-6
    lim = 10.0 * (atol + rtol * np.maximum(np.abs(sold[7:]), np.abs(ref_snew[7:])))
    if do_update:
        assert (np.abs(got_snew[7:] - ref_snew[7:]) / lim).max() <= 1.0
        assert (np.abs(got_snew[6] - ref_snew[6]) / ref_snew[6]).max() <= 1e-6
        for comp in (0, 1, 2, 3, 4, 5):  # rho, momenta, rho*E, rho*e: arithmetic on the inputs (rho: sum of the species)
            sc = np.abs(ref_snew[comp]).max()
            assert np.abs(got_snew[comp] - ref_snew[comp]).max() <= 1e-9 * sc, comp
    else:
        assert np.array_equal(got_snew, ref_snew)
    # I_R_k = (rhoY_new - rhoY_old)/dt - src_k inherits the species bar divided by dt
    assert (np.abs(got_IR[:N] - ref_IR[:N]) / (lim / DT)).max() <= 1.0
    # I_R_E = (rho*E_new - rho*E_old)/dt - src_E: a difference of O(rho*E) numbers, 1e-9 relative on rho*E
    assert np.abs(got_IR[N] - ref_IR[N]).max() <= 1e-9 * np.abs(sold[4]).max() / DT
    # heat release = - sum_k h_k I_R_k at the new temperature
    hi = np.abs(orc.thermo(ref_snew[6], (ref_snew[7:] / ref_snew[0]).T.copy())[3]).T  # [N, ncells]
    bound = (hi * lim / DT).sum(axis=0) + 1e-9 * np.abs(ref_IR[N + 1])
    assert (np.abs(got_IR[N + 1] - ref_IR[N + 1]) / bound).max() <= 1.0


@pytest.mark.parametrize("mech", ["LiDryer", "drm19"])
@pytest.mark.parametrize("do_update", [True, False])
def test_react_state_cvode_matches_restated_react_cpp(mech, do_update):
    orc = Oracle(mech)
    dev = torch.device("cuda:0")
    shape, pad, begin = (6, 8, 2), (2, 1, 1), (-2, 3, 7)
    sold, snew, nonrs = make_state(orc, shape, 42)
    if not do_update:  # react_init: S_old is S_new itself, sources zero (React.cpp:58-62, 141-142)
        nonrs = np.zeros_like(nonrs)
        snew = sold.copy()
    ref_snew, ref_IR, ref_FC = react_state_oracle(orc, sold, snew, nonrs, DT, do_update, "cvode", **CVO)
    N = orc.N
    r = ReactorB200.create("ReactorCvode", mech, **CV)
    r.init(1, 1)
    t_old, t_new, t_src = (to_fab(x, shape, dev, pad, fill=-1.0) for x in (sold, snew, nonrs))
    if not do_update:
        t_old = t_new
    t_IR = torch.full((N + 2,) + tuple(t_new.shape[1:]), -3.0, dtype=torch.float64, device=dev)
    t_w = torch.full((1,) + tuple(t_new.shape[1:]), -5.0, dtype=torch.float64, device=dev)
    new0 = t_new.clone()
    lo = (begin[0] + pad[0], begin[1] + pad[1], begin[2] + pad[2])
    hi = (lo[0] + shape[0] - 1, lo[1] + shape[1] - 1, lo[2] + shape[2] - 1)
    A = lambda t: Array4(t, begin)
    r.react_state(Box(lo, hi), A(t_old), A(t_new), A(t_src), A(t_IR), DT, do_update=do_update, work=A(t_w))
    torch.cuda.synchronize()
    got_snew, got_IR = from_fab(t_new, shape, pad), from_fab(t_IR, shape, pad)
    check_against_oracle(orc, sold, got_snew, got_IR, ref_snew, ref_IR, do_update)
    fc = from_fab(t_w, shape, pad)[0]
    assert np.all(np.abs(fc - ref_FC) <= np.maximum(3, 0.25 * ref_FC))
    # nothing outside the box is touched
    inside = torch.zeros(t_new.shape[1:], dtype=torch.bool, device=dev)
    inside[pad[2]:pad[2] + shape[2], pad[1]:pad[1] + shape[1], pad[0]:pad[0] + shape[0]] = True
    assert torch.equal(t_new[:, ~inside], new0[:, ~inside]) and torch.all(t_IR[:, ~inside] == -3.0) and torch.all(t_w[:, ~inside] == -5.0)
    assert r.kernel_name().startswith("bdf_kernel")
    r.close()


def test_react_state_rk64_and_unfused_path(monkeypatch):
    mech = "LiDryer"
    orc = Oracle(mech)
    dev = torch.device("cuda:0")
    shape = (16, 4, 2)
    sold, snew, nonrs = make_state(orc, shape, 7)
    N = orc.N
    A = lambda t: Array4(t)
    box = Box((0, 0, 0), (shape[0] - 1, shape[1] - 1, shape[2] - 1))
    # RK64 (fused)
    ref_snew, ref_IR, _ = react_state_oracle(orc, sold, snew, nonrs, DT, True, "rk64", atol=1e-10)
    r = ReactorB200.create("ReactorRK64", mech, ode__atol=1e-10)
    r.init(1, 1)
    t_old, t_new, t_src = (to_fab(x, shape, dev) for x in (sold, snew, nonrs))
    t_IR = torch.zeros((N + 2,) + tuple(t_new.shape[1:]), dtype=torch.float64, device=dev)
    r.react_state(box, A(t_old), A(t_new), A(t_src), A(t_IR), DT)
    torch.cuda.synchronize()
    got = from_fab(t_new, shape)
    rho = ref_snew[0]
    assert (np.abs(got[7:] - ref_snew[7:]) / np.maximum(np.abs(ref_snew[7:]), 1e-2 * rho)).max() <= 1e-10
    assert (np.abs(got[6] - ref_snew[6]) / ref_snew[6]).max() <= 1e-10
    gi = from_fab(t_IR, shape)
    assert np.abs(gi[N + 1] - ref_IR[N + 1]).max() <= 1e-6 * np.abs(ref_IR[N + 1]).max()
    r.close()
    # CVODE through the cooperative integrator: unfused pack / unpack kernels, same answer as the fused path
    ref_snew, ref_IR, _ = react_state_oracle(orc, sold, snew, nonrs, DT, True, "cvode", **CVO)
    monkeypatch.setenv("PRX_BDF_VARIANT", "8")
    r = ReactorB200.create("ReactorCvode", mech, **CV)
    r.init(1, 1)
    assert r.kernel_name().startswith("bdfc_kernel")
    t_old, t_new, t_src = (to_fab(x, shape, dev) for x in (sold, snew, nonrs))
    t_IR = torch.zeros((N + 2,) + tuple(t_new.shape[1:]), dtype=torch.float64, device=dev)
    old0, src0 = t_old.clone(), t_src.clone()
    r.react_state(box, A(t_old), A(t_new), A(t_src), A(t_IR), DT)
    torch.cuda.synchronize()
    assert torch.equal(t_old, old0) and torch.equal(t_src, src0)  # inputs are read only (rYsrc aliases the source FAB)
    check_against_oracle(orc, sold, from_fab(t_new, shape), from_fab(t_IR, shape), ref_snew, ref_IR, True)
    r.close()


@pytest.mark.parametrize("key", ["ReactorCvode", "ReactorRK64"])
def test_react_boxes_equals_box_by_box(key):
    """Several (ragged) boxes in one call == the same boxes one call at a time, bit for bit; also through react_state_boxes."""
    mech = "LiDryer"
    orc = Oracle(mech)
    dev = torch.device("cuda:0")
    N = orc.N
    shapes = [(8, 8, 8), (5, 3, 2), (16, 16, 4), (1, 1, 1), (7, 9, 3), (12, 4, 4), (8, 8, 8), (3, 3, 3), (4, 16, 2)]
    parm = dict(CV) if key == "ReactorCvode" else {"ode__atol": 1e-10}
    r = ReactorB200.create(key, mech, **parm)
    r.init(1, 1)

    def arrays(shape, seed):
        nx, ny, nz = shape
        st = reacteval_states(orc, shape)
        rng = np.random.default_rng(seed)
        st.rYsrc = 0.01 * rng.uniform(-1, 1, st.rY.shape) * st.rY / DT
        f = lambda a, nc: torch.from_numpy(np.ascontiguousarray(a).reshape(nc, nz, ny, nx).copy()).to(dev)
        return [f(st.rY, N), f(st.rYsrc, N), f(st.T, 1), f(st.rE, 1), f(st.rEsrc, 1),
                torch.zeros((1, nz, ny, nx), dtype=torch.float64, device=dev)]
    one = [arrays(s, i) for i, s in enumerate(shapes)]
    many = [[t.clone() for t in a] for a in one]
    boxes = [Box((0, 0, 0), (s[0] - 1, s[1] - 1, s[2] - 1)) for s in shapes]
    steps = 0.0
    for b, a in zip(boxes, one):
        r.react(b, *[Array4(t) for t in a], None, DT, 0.0)
        torch.cuda.synchronize()
        steps += float(a[5].sum())
    r.react_boxes(boxes, *[[Array4(a[i]) for a in many] for i in range(6)], None, DT, 0.0)
    torch.cuda.synchronize()
    for a, m in zip(one, many):
        for x, y in zip(a, m):
            assert torch.equal(x, y)
    if key == "ReactorRK64":  # react()'s return value: average steps per cell over the whole call
        ncell = sum(b.numPts() for b in boxes)
        assert abs(r.last_react_return() - int(steps / ncell)) <= 1
    r.close()


def test_react_state_boxes_equals_box_by_box():
    mech = "LiDryer"
    orc = Oracle(mech)
    dev = torch.device("cuda:0")
    N = orc.N
    shapes = [(8, 8, 4), (5, 3, 2), (16, 8, 2), (2, 2, 2), (8, 8, 4)]
    r = ReactorB200.create("ReactorCvode", mech, **CV)
    r.init(1, 1)
    sets = [make_state(orc, s, 100 + i) for i, s in enumerate(shapes)]
    boxes = [Box((0, 0, 0), (s[0] - 1, s[1] - 1, s[2] - 1)) for s in shapes]

    def fabs():
        out = []
        for (sold, snew, nonrs), s in zip(sets, shapes):
            t = [to_fab(x, s, dev) for x in (sold, snew, nonrs)]
            t.append(torch.zeros((N + 2,) + tuple(t[0].shape[1:]), dtype=torch.float64, device=dev))
            out.append(t)
        return out
    one, many = fabs(), fabs()
    for b, t in zip(boxes, one):
        r.react_state(b, *[Array4(x) for x in t], DT)
    torch.cuda.synchronize()
    r.react_state_boxes(boxes, *[[Array4(t[i]) for t in many] for i in range(4)], DT)
    torch.cuda.synchronize()
    for a, m in zip(one, many):
        for x, y in zip(a, m):
            assert torch.equal(x, y)
    r.close()


@pytest.mark.parametrize("mech", ["LiDryer", "drm19", "dodecane_lu"])
@pytest.mark.parametrize("layout", ["ragged", "aligned"])
def test_compute_temp_matches_oracle_rey2t(mech, layout):
    """pc_cmpTemp (PC/Source/Utilities.H:70-84): T <- REY2T(rho, e, Y) in place, on a box inside a padded, offset FAB: a ragged
    one (odd row length, misaligned rows: scalar loads) and an even, 16-byte aligned one (the 128-bit load path)."""
    orc = Oracle(mech)
    dev = torch.device("cuda:0")
    shape, pad, begin = ((37, 5, 3), (1, 2, 0), (4, -3, 0)) if layout == "ragged" else ((64, 4, 2), (2, 1, 1), (-6, 0, 3))
    # (the 128-bit form is selected by PRX_CT_VEC2=1, read once per process: covered when the suite runs with it set,
    #  tools/gpu_r02p.sh; results are identical either way)
    sold, _, _ = make_state(orc, shape, 3)
    rng = np.random.default_rng(5)
    sold[6] += rng.uniform(-150, 150, sold.shape[1])  # a poor guess
    t = to_fab(sold, shape, dev, pad, fill=-1.0)
    t0 = t.clone()
    r = ReactorB200.create("ReactorRK64", mech)
    r.init(1, 1)
    lo = (begin[0] + pad[0], begin[1] + pad[1], begin[2] + pad[2])
    hi = (lo[0] + shape[0] - 1, lo[1] + shape[1] - 1, lo[2] + shape[2] - 1)
    r.compute_temp([Box(lo, hi)], [Array4(t, begin)])
    torch.cuda.synchronize()
    got = from_fab(t, shape, pad)
    Y = (sold[7:] / sold[0]).T.copy()
    ref = orc.ey2t(sold[5] / sold[0], Y, sold[6])
    assert (np.abs(got[6] - ref) / ref).max() < 1e-12
    got[6] = sold[6]
    assert np.array_equal(got, sold)  # nothing but UTEMP changes
    inside = torch.zeros(t.shape[1:], dtype=torch.bool, device=dev)
    inside[pad[2]:pad[2] + shape[2], pad[1]:pad[1] + shape[1], pad[0]:pad[0] + shape[0]] = True
    assert torch.equal(t[:, ~inside], t0[:, ~inside])
    r.close()
