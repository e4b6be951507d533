"""Drop-in proof: the UNMODIFIED reference (PelePhysics ReactorBase + Factory, AMReX, SUNDIALS) linked with
integration/ReactorB200.cpp (the shim a maintainer adds) and libpelerx_b200.so.  The reference's own factory
creates "ReactorCvodeB200"/"ReactorRK64B200" by name and the reference's own virtual
ReactorBase::react(box, Array4..., dt, time) drives the CUDA kernels; results are compared with the reference's
CPU reactors created through the very same factory in the very same process.

integration/_build/libpele_b200_<mech>.so is built by integration/Makefile where /root/reference exists and
travels to the GPU box as a prebuilt file; nothing here reads /root/reference."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

from util import CVODE_SETTINGS, ROOT, assert_cvode_parity, assert_rk64_parity, golden, golden_state

dp = C.POINTER(C.c_double)


def lib_path(mech):
    return os.path.join(ROOT, "integration", "_build", "libpele_b200_%s.so" % mech)


def _p(a):
    return a.ctypes.data_as(dp)


def test_reference_factory_lists_b200_reactors():
    """No GPU needed: ask the reference factory for an unknown key; it prints its table and aborts."""
    mech = "LiDryer"
    if not os.path.exists(lib_path(mech)):
        pytest.skip("integration/_build not built (needs /root/reference at build time)")
    code = ("import ctypes as C\nL=C.CDLL(%r)\nL.peleref_init()\nL.peleref_reactor_create.restype=C.c_void_p\n"
            "L.peleref_reactor_create.argtypes=[C.c_char_p,C.c_int,C.c_int]\nL.peleref_reactor_create(b'NoSuchReactor',1,1)\n"
            % lib_path(mech))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert r.returncode != 0
    for key in ("ReactorCvode", "ReactorRK64", "ReactorCvodeB200", "ReactorRK64B200"):
        assert (" - %s\n" % key) in r.stdout, r.stdout[-600:]


class RefProcess:
    """The combined library, driven through oracle/ref/peleref_capi.cpp's extern "C" door."""

    def __init__(self, mech):
        self.L = L = C.CDLL(lib_path(mech))
        L.peleref_init()
        L.peleref_reactor_create.restype = C.c_void_p
        L.peleref_reactor_create.argtypes = [C.c_char_p, C.c_int, C.c_int]
        L.peleref_reactor_destroy.argtypes = [C.c_void_p]
        L.peleref_reactor_set_typ_vals.argtypes = [C.c_void_p, dp, C.c_int]
        L.peleref_set_parm.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p]
        L.peleref_react_box.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, dp, dp, dp, dp, dp, dp,
                                        C.POINTER(C.c_int), C.c_double, C.c_double]
        L.peleref_react_flat.argtypes = [C.c_void_p, dp, dp, dp, dp, C.c_double, C.c_double, C.c_int]

    def parm(self, **kw):
        for k, v in kw.items():
            pre, name = k.split("__")
            self.L.peleref_set_parm(pre.encode(), name.encode(), str(v).encode())

    def react(self, key, st, dt, nx=None, typ=None):
        h = self.L.peleref_reactor_create(key.encode(), 1, 1)
        if typ is not None:
            t = np.ascontiguousarray(typ, dtype=np.float64)
            self.L.peleref_reactor_set_typ_vals(h, _p(t), len(t))
        nx = st.ncells if nx is None else nx
        ny = st.ncells // nx
        rc = self.L.peleref_react_box(h, nx, ny, 1, _p(st.rY), _p(st.rYsrc), _p(st.T), _p(st.rE), _p(st.rEsrc), _p(st.FC),
                                      st.mask.ctypes.data_as(C.POINTER(C.c_int)), dt, 0.0)
        self.L.peleref_reactor_destroy(h)
        return rc


@pytest.mark.gpu
@pytest.mark.parametrize("mech", ["LiDryer", "drm19", "dodecane_lu"])
def test_b200_reactor_through_reference_factory(mech):
    if not os.path.exists(lib_path(mech)):
        pytest.skip("integration/_build/libpele_b200_%s.so not built" % mech)
    # run in a fresh interpreter: the reference library owns global AMReX state
    code = "import sys; sys.path[:0]=[%r,%r]; import test_dropin_reference as t; t._run(%r)" % (
        ROOT, os.path.join(ROOT, "tests"), mech)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, (r.stdout[-1500:], r.stderr[-1500:])
    assert "DROPIN-OK" in r.stdout


def _run(mech):
    g = golden(mech)
    dt = float(g["dt"])
    P = RefProcess(mech)
    # --- ReactorCvode (CPU, reference) vs ReactorCvodeB200 (GPU) through the same factory, same ParmParse keys
    for i, kw in enumerate(CVODE_SETTINGS[:2]):
        P.parm(ode__rtol=kw["rtol"], ode__atol=kw["atol"], cvode__max_order=kw["max_order"], cvode__solve_type=kw["solve_type"],
               ode__verbose=0)
        a, b = golden_state(g), golden_state(g)
        a.mask[1::5] = -1
        b.mask[1::5] = -1
        P.react("ReactorCvode", a, dt, nx=a.ncells // 2)
        rc = P.react("ReactorCvodeB200", b, dt, nx=a.ncells // 2)
        assert rc == 20
        assert_cvode_parity(b.rY, b.T, a.rY, a.T, kw["atol"], kw["rtol"])
        assert np.array_equal(a.rE, b.rE)
        skipped = a.mask == -1
        assert np.all(b.FC[skipped] == 0) and np.array_equal(b.rY[:, skipped], g["rY0"][:, skipped])
        assert abs(a.FC.sum() - b.FC.sum()) <= 0.25 * a.FC.sum()
    # typical values travel through ReactorBase::set_typ_vals_ode (non-virtual, read lazily by the shim)
    typ = np.concatenate([np.maximum(g["rY0"].max(axis=1), 1e-10), [2000.0]])
    a, b = golden_state(g), golden_state(g)
    P.react("ReactorCvode", a, dt, typ=typ)
    P.react("ReactorCvodeB200", b, dt, typ=typ)
    kw = CVODE_SETTINGS[1]
    lim = 10 * (typ[:-1, None] * kw["atol"] + kw["rtol"] * np.abs(a.rY))
    assert (np.abs(a.rY - b.rY) / lim).max() <= 1.0
    # --- RK64
    if "rk64_rY" in g:
        P.parm(ode__atol=1e-10)
        a, b = golden_state(g), golden_state(g)
        ra = P.react("ReactorRK64", a, dt)
        rb = P.react("ReactorRK64B200", b, dt)
        assert_rk64_parity(b.rY, b.T, a.rY, a.T)
        assert abs(ra - rb) <= 1 and np.array_equal(a.rE, b.rE)
        # flat (1-D) react, host pointers (ReactorBase.H:42-54)
        N = a.N
        s = golden_state(g)
        y1 = np.concatenate([s.rY.T, s.T[:, None]], axis=1).copy()
        y2 = y1.copy()
        e1, e2 = s.rE.copy(), s.rE.copy()
        ys, es = np.ascontiguousarray(s.rYsrc.T), s.rEsrc.copy()
        for key, y, e in (("ReactorRK64", y1, e1), ("ReactorRK64B200", y2, e2)):
            h = P.L.peleref_reactor_create(key.encode(), 1, 1)
            P.L.peleref_react_flat(h, _p(y), _p(ys), _p(e), _p(es), dt, 0.0, s.ncells)
            P.L.peleref_reactor_destroy(h)
        assert_rk64_parity(y2[:, :N].T, y2[:, N], y1[:, :N].T, y1[:, N])
        assert np.array_equal(e1, e2)
    print("DROPIN-OK", mech)
