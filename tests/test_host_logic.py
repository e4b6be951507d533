"""CPU tests of the PRODUCT's host-side pieces that do not need a GPU:
 - the C-ABI library loads and exports every symbol include/pelerx_b200.h declares;
 - the mechanism generator is deterministic and complete;
 - the resumable integrator state machines (pelec_b200/csrc/prx_integrators.cuh), instantiated
   for the host by tests/host_harness, reproduce the oracle bit-for-bit (logic check without a GPU);
 - argument checking / error codes of the entry points that return before touching the device."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from util import CVODE_SETTINGS, MECHS, ROOT, golden, golden_state

from oracle.pyoracle import Oracle, _p, ip

LIB = os.path.join(ROOT, "pelec_b200", "lib", "libpelerx_b200.so")


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "pelerx_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(prx_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    if not os.path.exists(LIB):
        pytest.fail("libpelerx_b200.so not built: run __graft_entry__.build()")
    lib = C.CDLL(LIB)
    syms = declared_symbols()
    assert len(syms) >= 18
    for s in syms:
        assert hasattr(lib, s), "missing export " + s
    # and nothing else leaks out of the library
    out = subprocess.check_output(["nm", "-D", "--defined-only", LIB]).decode()
    exported = [l.split()[-1] for l in out.splitlines() if " T " in l]
    assert all(e.startswith("prx_") for e in exported), exported
    # nvcc compiles a host call of a __device__-only function (reachable through a template) into a silent
    # exit(1): the library must not import exit at all
    und = subprocess.check_output(["nm", "-D", "--undefined-only", LIB]).decode()
    assert not re.search(r"\bU exit\b", und), "host code calls a __device__ function (nvcc emitted exit(1))"


def test_metadata_and_argument_checks_without_device():
    from pelec_b200 import reactor as R
    lib = R.load_library()
    assert lib.prx_num_species(0) == 9 and lib.prx_num_species(1) == 21 and lib.prx_num_species(2) == 53
    assert lib.prx_num_reactions(0) == 21 and lib.prx_num_reactions(1) == 84 and lib.prx_num_reactions(2) == 268
    assert lib.prx_mechanism_name(2) == b"dodecane_lu"
    assert lib.prx_num_species(7) == -1
    cfg = R.prx_config()
    assert lib.prx_default_config(C.byref(cfg)) == 0
    assert (cfg.rtol, cfg.atol, cfg.max_order, cfg.max_nls_iters) == (1e-6, 1e-10, 2, 3)
    h = C.c_void_p()
    bad = R.prx_config()
    lib.prx_default_config(C.byref(bad))
    bad.reactor_type = 3  # reference: amrex::Abort("Wrong reactor type...")
    assert lib.prx_create(C.byref(bad), C.byref(h)) == -1
    bad.reactor_type = 1
    bad.solve_type = 9
    assert lib.prx_create(C.byref(bad), C.byref(h)) == -4
    bad.solve_type = 0
    bad.mechanism = 11
    assert lib.prx_create(C.byref(bad), C.byref(h)) == -1
    import torch
    if not torch.cuda.is_available():
        # no CPU fallback: creating a reactor without a device must fail loudly
        assert lib.prx_create(C.byref(cfg), C.byref(h)) == -2
        with pytest.raises(R.PrxError):
            r = R.ReactorB200.create("ReactorCvode", "LiDryer")
            r.init(1, 1)
    with pytest.raises(KeyError):
        R.ReactorB200.create("ReactorFoo")
    r = R.ReactorB200.create("ReactorCvode", "LiDryer", cvode__solve_type="GMRES")
    with pytest.raises(R.PrxError):
        r.init(1, 1)


def test_generator_is_deterministic():
    from pelec_b200.mechgen import gen
    for m in MECHS:
        a, b = gen.emit(m, "cuda"), gen.emit(m, "cuda")
        assert a == b and "production_rate" in a and "ajacobian" in a
        d = gen.load(m)
        assert a.count("{ // R") == 2 * len(d["reactions"])


@pytest.fixture(scope="module")
def harness():
    hdir = os.path.join(ROOT, "tests", "host_harness")
    libs = {}
    for m in MECHS:
        so = os.path.join(hdir, "libhh_%s.so" % m)
        src = os.path.join(hdir, "harness.cpp")
        deps = [src] + [os.path.join(ROOT, "pelec_b200", "csrc", f) for f in
                        ("prx_integrators.cuh", "prx_chem.cuh", "prx_tchem.cuh", "gen/mech_%s.cuh" % m)]
        if not os.path.exists(deps[-1]):
            from pelec_b200.mechgen import gen
            gen.generate_all(ROOT, (m,))
        if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
            subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-std=c++17",
                                   "-DPRX_MECH_HEADER=\"gen/mech_%s.cuh\"" % m,
                                   "-I" + os.path.join(ROOT, "pelec_b200", "csrc"), "-o", so, src])
        libs[m] = C.CDLL(so)
    return libs


@pytest.mark.parametrize("mech", MECHS)
def test_state_machines_reproduce_oracle(harness, mech):
    hh = harness[mech]
    orc = Oracle(mech)
    g = golden(mech)
    dt = float(g["dt"])
    for kw in CVODE_SETTINGS:
        a, b = golden_state(g), golden_state(g)
        a.mask[1::4] = -1
        b.mask[1::4] = -1
        so = orc.react_cvode(b, dt, want_stats=True, **kw)
        stats = np.zeros((a.ncells, 8), dtype=np.int64)
        hh.hh_react_cvode(C.c_int(1), C.c_double(kw["rtol"]), C.c_double(kw["atol"]), None, kw["max_order"], 3,
                          1 if kw["solve_type"] == "denseAJ_direct" else 0, a.ncells, _p(a.rY), _p(a.rYsrc),
                          _p(a.T), _p(a.rE), _p(a.rEsrc), _p(a.FC), a.mask.ctypes.data_as(ip), C.c_double(dt),
                          C.c_double(0.0), stats.ctypes.data_as(C.POINTER(C.c_long)))
        assert np.array_equal(a.rY, b.rY) and np.array_equal(a.T, b.T) and np.array_equal(a.FC, b.FC)
        live = np.where(a.mask != -1)[0]
        assert all(stats[i, 0] == so[i].nst and stats[i, 3] == so[i].nje and stats[i, 4] == so[i].nsetups for i in live)
    if "rk64_rY" in g:
        a, b = golden_state(g), golden_state(g)
        orc.react_rk64(b, dt, atol=1e-10)
        hh.hh_react_rk64(C.c_int(1), C.c_double(1e-10), 10, 5, 500, a.ncells, _p(a.rY), _p(a.rYsrc), _p(a.T),
                         _p(a.rE), _p(a.rEsrc), _p(a.FC), C.c_double(dt), C.c_double(0.0))
        assert np.array_equal(a.rY, b.rY) and np.array_equal(a.T, b.T) and np.array_equal(a.FC, b.FC)


@pytest.mark.parametrize("mech", MECHS)
def test_table_driven_chemistry_and_column_store_are_bit_identical(harness, mech):
    """The table-driven RHS (prx_tchem.cuh: the mechanism as data, both table address spaces) and the
    strided shared-memory-column store of the integrator state must not change a single bit with respect
    to the generated straight-line code / thread-local arrays (which the test above pins to the oracle)."""
    hh = harness[mech]
    g = golden(mech)
    dt = float(g["dt"])
    st = golden_state(g)
    n, N = st.ncells, st.rY.shape[0]
    rng = np.random.default_rng(7)
    # unit level: RHS on the golden states, with external sources and at a later time
    y = np.concatenate([st.rY.T, st.T[:, None]], axis=1).copy()
    rYsrc = (rng.standard_normal((n, N)) * 1e-3 * st.rY.T).copy()
    rEsrc = (rng.standard_normal(n) * 1e-3 * np.abs(st.rE) / dt).copy()
    outs = []
    for tab in (0, 1):
        yd = np.zeros_like(y)
        hh.hh_rhs(tab, n, C.c_double(0.3 * dt), 1, _p(y), _p(st.rE), _p(rEsrc), _p(rYsrc), _p(yd))
        outs.append(yd)
    assert np.array_equal(outs[0], outs[1])
    assert np.abs(outs[0]).max() > 0
    # integrator level
    for kw in CVODE_SETTINGS[:2] if mech == "dodecane_lu" else CVODE_SETTINGS:
        res = []
        for variant in (0, 1, 2):
            a = golden_state(g)
            a.mask[1::4] = -1
            hh.hh_react_cvode_variant(variant, 1, C.c_double(kw["rtol"]), C.c_double(kw["atol"]), None, kw["max_order"], 3,
                                      1 if kw["solve_type"] == "denseAJ_direct" else 0, a.ncells, _p(a.rY), _p(a.rYsrc),
                                      _p(a.T), _p(a.rE), _p(a.rEsrc), _p(a.FC), a.mask.ctypes.data_as(ip),
                                      C.c_double(dt), C.c_double(0.0), None)
            res.append(a)
        for b in res[1:]:
            assert np.array_equal(res[0].rY, b.rY) and np.array_equal(res[0].T, b.T) and np.array_equal(res[0].FC, b.FC)
    if "rk64_rY" in g:
        res = []
        for tab in (0, 1):
            a = golden_state(g)
            hh.hh_react_rk64_variant(tab, 1, C.c_double(1e-10), 10, 5, 500, a.ncells, _p(a.rY), _p(a.rYsrc), _p(a.T),
                                     _p(a.rE), _p(a.rEsrc), _p(a.FC), C.c_double(dt), C.c_double(0.0))
            res.append(a)
        assert np.array_equal(res[0].rY, res[1].rY) and np.array_equal(res[0].T, res[1].T)


@pytest.mark.parametrize("mech", MECHS)
@pytest.mark.parametrize("reactor_type", [1, 2])
def test_table_driven_jacobian_is_bit_identical(harness, mech, reactor_type):
    """ajacobian_tab (prx_tchem.cuh: per-reaction derivatives from the mechanism tables, incl. the reference's
    stale-dqdc quirk for third bodies of efficiency 0) against the generated straight-line aJacobian, on the
    golden states and on perturbed ones, for the constant-volume and the constant-pressure forms."""
    hh = harness[mech]
    g = golden(mech)
    st = golden_state(g)
    n = st.ncells
    rng = np.random.default_rng(11)
    y = np.concatenate([st.rY.T, st.T[:, None]], axis=1)
    y = np.concatenate([y, y * (1.0 + 0.3 * rng.random(y.shape))], axis=0).copy()
    y[n:, -1] = rng.uniform(400.0, 3000.0, n)
    NEQ = y.shape[1]
    out = []
    for tab in (0, 1):
        J = np.zeros((y.shape[0], NEQ, NEQ))
        hh.hh_jac(tab, y.shape[0], reactor_type, _p(y), _p(J))
        out.append(J)
    assert np.abs(out[0]).max() > 0 and np.all(np.isfinite(out[0]))
    assert np.array_equal(out[0], out[1])


@pytest.mark.parametrize("mech", MECHS[:2])
def test_register_staged_lu_is_bit_identical(harness, mech):
    """dense_getrf_small / dense_getrs_small (loads of an elimination step hoisted above its stores) against the
    reference-order loops: random systems, systems with tied pivots, zero entries and singular columns."""
    hh = harness[mech]
    NEQ = hh.hh_num_species() + 1
    rng = np.random.default_rng(3)
    n = 64
    A = rng.standard_normal((n, NEQ, NEQ))
    A[1::4] = np.round(A[1::4] * 2.0)            # ties and exact zeros
    A[2::8, 3, :] = 0.0                           # a zero column (column-major: [col, row]) -> singular
    A[5::8] += np.eye(NEQ) * 50.0
    b = rng.standard_normal((n, NEQ))
    res = []
    for small in (0, 1, 2):
        a, x = A.copy(), b.copy()
        piv = np.zeros((n, NEQ), dtype=np.int32)
        rc = hh.hh_lu(small, n, _p(a), piv.ctypes.data_as(ip), _p(x))
        res.append((rc, a, piv, x))
    for other in res[1:]:
        assert res[0][0] == other[0] and res[0][0] > 0
        for u, v in zip(res[0][1:], other[1:]):
            assert np.array_equal(u, v, equal_nan=True)
