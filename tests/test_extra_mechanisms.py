"""Mechanisms beyond the three of BASELINE.json ("the other 41 mechanisms"): the pipeline
tools/extract_mechanism.py -> pelec_b200/mechgen/gen.py -> one kernel translation unit per mechanism is not specific to
LiDryer / drm19 / dodecane_lu.  pelec_b200/csrc/Makefile EXTRA_MECHS compiles BurkeDryer (13 species, 27 reactions) and
FFCM1_Red (21 species, 124 reactions) into the default library; tools/validate_mechanisms.py checks the generated code of every
reacting non-QSS mechanism of the reference against the reference library (profiles/mechanism_coverage.json).

 * CPU: the generated C oracle of the extra mechanisms reproduces the reference library bit for bit (RHS, transport) where that
   library was built (oracle/ref/Makefile MECHS=...); the C-ABI registry resolves them by name;
 * GPU: RHS / Jacobian probes, ReactorCvode, ReactorRK64 and the transport kernel against the oracle.

Nothing here reads /root/reference."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle.pyoracle import Oracle, Reference, reacteval_states
from util import assert_cvode_parity, assert_rk64_parity, have_reference

EXTRA = ["BurkeDryer", "FFCM1_Red"]
DT = 1e-6


def states(orc, n=64):
    st = reacteval_states(orc, (1, n, 1))
    rho = st.rY.sum(axis=0)
    return st, rho


@pytest.mark.parametrize("mech", EXTRA)
def test_generated_oracle_reproduces_reference_library(mech):
    if not have_reference(mech):
        pytest.skip("oracle/_ref/libpeleref_%s.so not built (make -C oracle/ref MECHS=%s)" % (mech, mech))
    o, r = Oracle(mech), Reference(mech)
    st, rho = states(o)
    N = o.N
    y = np.concatenate([st.rY.T, st.T[:, None]], axis=1).copy()
    z = np.zeros(st.ncells)
    fa = o.fkernelspec(0.0, y, st.rE, z, np.zeros((st.ncells, N)))
    fb = r.fkernelspec(0.0, y, st.rE, z, np.zeros((st.ncells, N)))
    assert np.array_equal(fa, fb)
    Y = (st.rY / rho).T.copy()
    for a, b in zip(o.transport(st.T, rho, Y), r.transport(st.T, rho, Y)):
        assert np.array_equal(a, b)
    # one CVODE call and one RK64 call: the integrators are mechanism-independent, the chemistry is not
    a, b = st.copy(), st.copy()
    o.react_rk64(a, DT, atol=1e-10)
    r.react_rk64(b, DT, atol=1e-10)
    assert np.array_equal(a.rY, b.rY) and np.array_equal(a.T, b.T)
    a, b = st.copy(), st.copy()
    o.react_cvode(a, DT, rtol=1e-8, atol=1e-10, max_order=3)
    r.react_cvode(b, DT, rtol=1e-8, atol=1e-10, max_order=3)
    assert_cvode_parity(a.rY, a.T, b.rY, b.T, 1e-10, 1e-8)


def test_library_registry_lists_the_extra_mechanisms():
    from pelec_b200.reactor import LIB_PATH
    lib = C.CDLL(LIB_PATH)  # symbols only: no device needed
    lib.prx_mechanism_name.restype = C.c_char_p
    lib.prx_mechanism_id.argtypes = [C.c_char_p]
    names = [lib.prx_mechanism_name(i).decode() for i in range(lib.prx_num_mechanisms())]
    assert names[:3] == ["LiDryer", "drm19", "dodecane_lu"]  # ABI constants
    for m in EXTRA:
        i = lib.prx_mechanism_id(m.encode())
        assert i >= 3 and names[i] == m and lib.prx_num_species(i) == Oracle(m).N
    assert lib.prx_mechanism_id(b"no_such_mechanism") < 0


@pytest.mark.gpu
@pytest.mark.parametrize("mech", EXTRA)
def test_gpu_extra_mechanism_matches_oracle(mech):
    torch = pytest.importorskip("torch")
    from pelec_b200.reactor import Array4, Box, ReactorB200
    orc = Oracle(mech)
    st, rho = states(orc, 96)
    N, n = orc.N, st.ncells
    # RHS and Jacobian probes
    r = ReactorB200.create("ReactorCvode", mech, ode__rtol=1e-8, ode__atol=1e-10, cvode__max_order=3)
    r.init(1, 1)
    y = np.concatenate([st.rY.T, st.T[:, None]], axis=1).copy()
    z = np.zeros(n)
    yd = r.probe_rhs(0.0, y, st.rE, z, np.zeros((n, N)))
    ref = orc.fkernelspec(0.0, y, st.rE, z, np.zeros((n, N)))
    sc = np.abs(ref[:, :-1]).max(axis=1, keepdims=True)
    assert (np.abs(yd[:, :-1] - ref[:, :-1]) / sc).max() < 1e-11
    assert (np.abs(yd[:, -1] - ref[:, -1]) / np.abs(ref[:, -1]).max()).max() < 1e-9
    J = r.probe_jac(y)
    Jo = np.zeros_like(J)
    dp = C.POINTER(C.c_double)
    orc.lib.orc_cjac(n, 1, y.ctypes.data_as(dp), Jo.ctypes.data_as(dp))
    assert (np.abs(J - Jo) / np.abs(Jo).max(axis=(1, 2), keepdims=True)).max() < 1e-11
    # ReactorCvode
    a, b = st.copy(), st.copy()
    r.react_host(n, 1, 1, a.rY, a.rYsrc, a.T, a.rE, a.rEsrc, a.FC, a.mask, DT, 0.0)
    orc.react_cvode(b, DT, rtol=1e-8, atol=1e-10, max_order=3)
    assert_cvode_parity(a.rY, a.T, b.rY, b.T, 1e-10, 1e-8)
    assert r.stats()["n_failed"] == 0
    r.close()
    # ReactorRK64
    r = ReactorB200.create("ReactorRK64", mech, ode__atol=1e-10)
    r.init(1, 1)
    a, b = st.copy(), st.copy()
    r.react_host(n, 1, 1, a.rY, a.rYsrc, a.T, a.rE, a.rEsrc, a.FC, a.mask, DT, 0.0)
    orc.react_rk64(b, DT, atol=1e-10)
    assert_rk64_parity(a.rY, a.T, b.rY, b.T)
    # transport kernel
    dev = torch.device("cuda:0")
    q = torch.zeros((8 + N, 1, 1, n), dtype=torch.float64, device=dev)
    q[0, 0, 0] = torch.from_numpy(rho).to(dev)
    q[7, 0, 0] = torch.from_numpy(st.T).to(dev)
    q[8:, 0, 0] = torch.from_numpy(np.ascontiguousarray(st.rY / rho)).to(dev)
    coef = torch.zeros((N + 3, 1, 1, n), dtype=torch.float64, device=dev)
    r.transport([Box((0, 0, 0), (n - 1, 0, 0))], [Array4(q[8:])], [Array4(q[7:8])], [Array4(q[0:1])], [Array4(coef[:N])],
                [Array4(coef[N:N + 1])], [Array4(coef[N + 1:N + 2])], [Array4(coef[N + 2:N + 3])])
    torch.cuda.synchronize()
    got = coef.reshape(N + 3, -1).cpu().numpy()
    D, mu, xi, lam = orc.transport(st.T, rho, (st.rY / rho).T.copy())
    rel = lambda x, y: float(np.max(np.abs(x - y) / np.abs(y)))
    assert rel(got[:N].T, D) <= 1e-12 and rel(got[N], mu) <= 1e-12 and rel(got[N + 1], xi) <= 1e-12 and rel(got[N + 2], lam) <= 1e-12
    r.close()
