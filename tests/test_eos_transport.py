"""The pointwise kernels either side of the reactor: pc_ctoprim (PC/Source/Utilities.H:242-321) and
SimpleTransport::get_transport_coeffs (PP/Source/Transport/Simple.H:230-447), against fixtures made by the UNMODIFIED
reference (tests/golden/make_golden4.py: the PeleC application library for ctoprim, the PelePhysics library for transport).

 * CPU (-m "not gpu"): the C restatement (oracle/chem_oracle.c orc_ctoprim / orc_transport) reproduces the fixtures BIT FOR BIT;
 * GPU (-m gpu): prx_ctoprim_boxes / prx_transport_boxes through the C ABI - on the fixture states, and on boxes inside padded,
   offset FABs (several boxes per call) against the pinned oracle.

Bars on the device: T from the Newton iteration 1e-12 relative (its own tolerance is 1e-6 K, but it converges quadratically
to rounding); every other output 1e-12 relative - the device's exp / log / cbrt differ from glibc's in the last bit.
Nothing here reads /root/reference."""
import os

import numpy as np
import pytest

from oracle.pyoracle import Oracle, reacteval_states

GOLD = os.path.join(os.path.dirname(__file__), "golden")
MECHS = ["LiDryer", "drm19", "dodecane_lu"]


def gold(name):
    z = np.load(os.path.join(GOLD, name))
    return {k: z[k] for k in z.files}


def test_ctoprim_oracle_is_bit_exact_with_pelec():
    g = gold("ctoprim_pmf_LiDryer.npz")
    q, qa = Oracle("LiDryer").ctoprim(g["u"])
    assert np.array_equal(q, g["q"]) and np.array_equal(qa, g["qa"])
    assert (g["q"][8 + 3] == 0.0).sum() >= 80  # the flushed trace mass fractions are part of the fixture


@pytest.mark.parametrize("mech", MECHS)
def test_transport_oracle_is_bit_exact_with_the_reference(mech):
    g = gold("transport_%s.npz" % mech)
    D, mu, xi, lam = Oracle(mech).transport(g["T"], g["rho"], g["Y"])
    assert np.array_equal(D, g["D"]) and np.array_equal(mu, g["mu"]) and np.array_equal(xi, g["xi"]) and np.array_equal(lam, g["lam"])


# ---------------------------------------------------------------------------------------------- GPU
def _fab(torch, a, shape, dev, pad, fill):
    nx, ny, nz = shape
    px, py, pz = pad
    t = torch.full((a.shape[0], nz + 2 * pz, ny + 2 * py, nx + 2 * px), fill, dtype=torch.float64, device=dev)
    t[:, pz:pz + nz, py:py + ny, px:px + nx] = torch.from_numpy(np.ascontiguousarray(a).reshape(a.shape[0], nz, ny, nx).copy()).to(dev)
    return t


def _unfab(t, shape, pad):
    nx, ny, nz = shape
    px, py, pz = pad
    return t[:, pz:pz + nz, py:py + ny, px:px + nx].reshape(t.shape[0], -1).cpu().numpy()


def _rel(a, b, floor=0.0):
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), floor)))


@pytest.mark.gpu
def test_gpu_ctoprim_matches_pelec_fixture():
    torch = pytest.importorskip("torch")
    from pelec_b200.reactor import Array4, Box, ReactorB200
    g = gold("ctoprim_pmf_LiDryer.npz")
    dev = torch.device("cuda:0")
    n = g["u"].shape[1]
    r = ReactorB200.create("ReactorRK64", "LiDryer")
    r.init(1, 1)
    u = _fab(torch, g["u"], (n, 1, 1), dev, (0, 0, 0), 0.0)
    q = torch.zeros((g["q"].shape[0], 1, 1, n), dtype=torch.float64, device=dev)
    qa = torch.zeros((6, 1, 1, n), dtype=torch.float64, device=dev)
    r.ctoprim([Box((0, 0, 0), (n - 1, 0, 0))], [Array4(u)], [Array4(q)], [Array4(qa)])
    torch.cuda.synchronize()
    gq, gqa = q.reshape(q.shape[0], -1).cpu().numpy(), qa.reshape(6, -1).cpu().numpy()
    assert np.array_equal(gq[:4], g["q"][:4]) and np.array_equal(gq[8:], g["q"][8:])  # rho, velocities, Y: pure arithmetic
    for c in (5, 6, 7):
        assert _rel(gq[c], g["q"][c]) <= 1e-12, c
    # QGAME = p / (rho e) + 1 with rho e < 0 in these cells (Fuego energies carry the formation enthalpy): 0.0013 = 1 - 0.9987,
    # a cancellation that amplifies the last-bit differences of p and e 800-fold; the bar is on p / (rho e) itself
    assert _rel(gq[4], g["q"][4], 1.0) <= 1e-12
    for c in range(6):
        assert _rel(gqa[c], g["qa"][c]) <= 1e-12, c
    r.close()


@pytest.mark.gpu
@pytest.mark.parametrize("mech", MECHS)
def test_gpu_transport_matches_reference_fixture(mech):
    torch = pytest.importorskip("torch")
    from pelec_b200.reactor import Array4, Box, ReactorB200
    g = gold("transport_%s.npz" % mech)
    dev = torch.device("cuda:0")
    n, N = g["T"].size, g["Y"].shape[1]
    r = ReactorB200.create("ReactorRK64", mech)
    r.init(1, 1)
    # PeleC hands component-offset views of ONE primitive FAB (q.array(QFS), q.array(QTEMP), q.array(QRHO)) and of ONE
    # coefficient FAB (dComp_rhoD, dComp_mu, dComp_xi, dComp_lambda): do the same
    q = torch.zeros((8 + N, 1, 1, n), dtype=torch.float64, device=dev)
    q[0, 0, 0] = torch.from_numpy(g["rho"]).to(dev)
    q[7, 0, 0] = torch.from_numpy(g["T"]).to(dev)
    q[8:, 0, 0] = torch.from_numpy(np.ascontiguousarray(g["Y"].T)).to(dev)
    coef = torch.full((N + 3, 1, 1, n), -1.0, dtype=torch.float64, device=dev)
    box = [Box((0, 0, 0), (n - 1, 0, 0))]
    r.transport(box, [Array4(q[8:])], [Array4(q[7:8])], [Array4(q[0:1])], [Array4(coef[:N])], [Array4(coef[N:N + 1])],
                [Array4(coef[N + 1:N + 2])], [Array4(coef[N + 2:N + 3])])
    torch.cuda.synchronize()
    got = coef.reshape(N + 3, -1).cpu().numpy()
    assert _rel(got[:N].T, g["D"]) <= 1e-12
    assert _rel(got[N], g["mu"]) <= 1e-12 and _rel(got[N + 1], g["xi"]) <= 1e-12 and _rel(got[N + 2], g["lam"]) <= 1e-12
    r.close()


@pytest.mark.gpu
@pytest.mark.parametrize("mech", MECHS)
def test_gpu_ctoprim_and_transport_on_padded_boxes_match_oracle(mech):
    torch = pytest.importorskip("torch")
    from pelec_b200.reactor import Array4, Box, ReactorB200
    orc = Oracle(mech)
    dev = torch.device("cuda:0")
    N = orc.N
    r = ReactorB200.create("ReactorRK64", mech)
    r.init(1, 1)
    rng = np.random.default_rng(17)
    boxes, us, qs, qas, coefs, refs, metas = [], [], [], [], [], [], []
    for shape, pad, begin in (((37, 5, 3), (1, 2, 0), (4, -3, 0)), ((16, 8, 2), (2, 1, 1), (-6, 0, 3))):
        st = reacteval_states(orc, shape)
        nc = st.ncells
        rho = st.rY.sum(axis=0)
        vel = rng.normal(size=(3, nc)) * 4.0e3
        u = np.zeros((7 + N, nc))
        u[0], u[1:4], u[5], u[6], u[7:] = rho, rho * vel, st.rE, st.T + rng.uniform(-80, 80, nc), st.rY
        u[4] = st.rE + 0.5 * rho * (vel ** 2).sum(axis=0)
        lo = tuple(begin[d] + pad[d] for d in range(3))
        hi = tuple(lo[d] + shape[d] - 1 for d in range(3))
        boxes.append(Box(lo, hi))
        us.append(Array4(_fab(torch, u, shape, dev, pad, -1.0), begin))
        qs.append(Array4(_fab(torch, np.zeros((8 + N, nc)), shape, dev, pad, -2.0), begin))
        qas.append(Array4(_fab(torch, np.zeros((6, nc)), shape, dev, pad, -3.0), begin))
        coefs.append(_fab(torch, np.zeros((N + 3, nc)), shape, dev, pad, -4.0))
        refs.append(orc.ctoprim(u))
        metas.append((shape, pad, begin))
    r.ctoprim(boxes, us, qs, qas)
    torch.cuda.synchronize()
    for b, (shape, pad, begin) in enumerate(metas):
        gq, gqa = _unfab(qs[b].t, shape, pad), _unfab(qas[b].t, shape, pad)
        rq, rqa = refs[b]
        floor = np.full((gq.shape[0], 1), 1e-300)
        floor[4] = 1.0  # QGAME = p / (rho e) + 1: see test_gpu_ctoprim_matches_pelec_fixture
        assert float(np.max(np.abs(gq - rq) / np.maximum(np.abs(rq), floor))) <= 1e-12 and _rel(gqa, rqa) <= 1e-12
        inside = torch.zeros(qs[b].t.shape[1:], dtype=torch.bool, device=dev)
        inside[pad[2]:pad[2] + shape[2], pad[1]:pad[1] + shape[1], pad[0]:pad[0] + shape[0]] = True
        assert torch.all(qs[b].t[:, ~inside] == -2.0) and torch.all(qas[b].t[:, ~inside] == -3.0)
    # transport from the primitives just computed (views of q, as PeleC's diffusion driver passes them)
    Yv = [Array4(q.t[8:], m[2]) for q, m in zip(qs, metas)]
    Tv = [Array4(q.t[7:8], m[2]) for q, m in zip(qs, metas)]
    Rv = [Array4(q.t[0:1], m[2]) for q, m in zip(qs, metas)]
    Dv = [Array4(c[:N], m[2]) for c, m in zip(coefs, metas)]
    Mv = [Array4(c[N:N + 1], m[2]) for c, m in zip(coefs, metas)]
    Xv = [Array4(c[N + 1:N + 2], m[2]) for c, m in zip(coefs, metas)]
    Lv = [Array4(c[N + 2:N + 3], m[2]) for c, m in zip(coefs, metas)]
    r.transport(boxes, Yv, Tv, Rv, Dv, Mv, Xv, Lv)
    torch.cuda.synchronize()
    for b, (shape, pad, begin) in enumerate(metas):
        gq = _unfab(qs[b].t, shape, pad)
        D, mu, xi, lam = orc.transport(gq[7], gq[0], np.ascontiguousarray(gq[8:].T))
        got = _unfab(coefs[b], shape, pad)
        assert _rel(got[:N].T, D) <= 1e-12 and _rel(got[N], mu) <= 1e-12 and _rel(got[N + 1], xi) <= 1e-12
        assert _rel(got[N + 2], lam) <= 1e-12
        inside = torch.zeros(coefs[b].shape[1:], dtype=torch.bool, device=dev)
        inside[pad[2]:pad[2] + shape[2], pad[1]:pad[1] + shape[1], pad[0]:pad[0] + shape[0]] = True
        assert torch.all(coefs[b][:, ~inside] == -4.0)
    r.close()
