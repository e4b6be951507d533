"""TEST INFRASTRUCTURE — ctypes doors onto (a) the CPU restatement ``oracle/_build/liboracle_<mech>.so``
(class ``Oracle``) and (b) the UNMODIFIED reference ``oracle/_ref/libpeleref_<mech>.so`` (class ``Reference``).

Only tests/, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs import this
module; the product (``pelec_b200``) never does.

Both classes expose the same methods so a test can run one body against either:
    react_rk64(state, dt, time=0, atol=...) / react_cvode(state, dt, rtol, atol, max_order, ...)
where ``state`` is a ``BoxState`` (SoA arrays laid out like an AMReX FAB: ``a[comp, cell]``).
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
MECHS = ("LiDryer", "drm19", "dodecane_lu")
FUEL = {"LiDryer": "H2", "drm19": "CH4", "dodecane_lu": "NC12H26", "BurkeDryer": "H2", "FFCM1_Red": "CH4"}
dp = C.POINTER(C.c_double)
ip = C.POINTER(C.c_int)


def _p(a):
    return a.ctypes.data_as(dp)


class BoxState:
    """Reactor-boundary arrays for ncells cells (the seven Array4s of ReactorBase::react)."""

    def __init__(self, rY, T, rE, rYsrc=None, rEsrc=None, mask=None):
        self.rY = np.ascontiguousarray(rY, dtype=np.float64)  # [N, ncells]
        self.N, self.ncells = self.rY.shape
        self.T = np.ascontiguousarray(T, dtype=np.float64).reshape(self.ncells)
        self.rE = np.ascontiguousarray(rE, dtype=np.float64).reshape(self.ncells)
        self.rYsrc = np.zeros_like(self.rY) if rYsrc is None else np.ascontiguousarray(rYsrc, dtype=np.float64)
        self.rEsrc = np.zeros(self.ncells) if rEsrc is None else np.ascontiguousarray(rEsrc, dtype=np.float64)
        self.mask = np.ones(self.ncells, dtype=np.int32) if mask is None else np.ascontiguousarray(mask, dtype=np.int32)
        self.FC = np.zeros(self.ncells)

    def copy(self):
        s = BoxState(self.rY.copy(), self.T.copy(), self.rE.copy(), self.rYsrc.copy(), self.rEsrc.copy(), self.mask.copy())
        s.FC = self.FC.copy()
        return s

    def take(self, idx):
        return BoxState(self.rY[:, idx], self.T[idx], self.rE[idx], self.rYsrc[:, idx], self.rEsrc[idx], self.mask[idx])


class _RK64Params(C.Structure):
    _fields_ = [("reactor_type", C.c_int), ("atol", C.c_double), ("nsubsteps_guess", C.c_int),
                ("nsubsteps_min", C.c_int), ("nsubsteps_max", C.c_int)]


def _cvode_params_type(neq):
    class P(C.Structure):
        _fields_ = [("reactor_type", C.c_int), ("rtol", C.c_double), ("atol", C.c_double),
                    ("typ_vals", C.c_double * neq), ("max_order", C.c_int), ("max_nls_iters", C.c_int),
                    ("analytic_jac", C.c_int), ("mxstep", C.c_int), ("maxnef", C.c_int), ("msbj", C.c_int),
                    ("clean_init_massfrac", C.c_int)]
    return P


class CvodeStats(C.Structure):
    _fields_ = [("nst", C.c_long), ("nfe", C.c_long), ("nfeDQ", C.c_long), ("nje", C.c_long),
                ("nsetups", C.c_long), ("nni", C.c_long), ("ncfn", C.c_long), ("netf", C.c_long),
                ("flag", C.c_int), ("qlast", C.c_int), ("hlast", C.c_double)]


def build_oracle(mechs=MECHS):
    """Compile oracle/chem_oracle.c for the given mechanisms (generating the C headers first)."""
    import sys
    root = os.path.dirname(HERE)
    sys.path.insert(0, root)
    from pelec_b200.mechgen import gen
    gen.generate_all(root, tuple(mechs))
    subprocess.check_call(["make", "-C", HERE, "-j4", "MECHS=" + " ".join(mechs)], stdout=subprocess.DEVNULL)


def build_reference(mechs=MECHS):
    """Compile the unmodified reference into oracle/_ref (needs /root/reference)."""
    subprocess.check_call(["make", "-C", os.path.join(HERE, "ref"), "-j8", "MECHS=" + " ".join(mechs)],
                          stdout=subprocess.DEVNULL)


class Oracle:
    """The CPU restatement (oracle/chem_oracle.c)."""
    kind = "port"

    def __init__(self, mech):
        path = os.path.join(HERE, "_build", "liboracle_%s.so" % mech)
        if not os.path.exists(path):
            build_oracle((mech,))
        self.lib = L = C.CDLL(path)
        self.mech = mech
        self.N = L.orc_num_species()
        self.neq = self.N + 1
        self.nreac = L.orc_num_reactions()
        L.orc_species_name.restype = C.c_char_p
        self.species = [L.orc_species_name(k).decode() for k in range(self.N)]
        self.CvodeParams = _cvode_params_type(self.neq)
        L.orc_react_rk64.argtypes = [C.POINTER(_RK64Params), C.c_int, dp, dp, dp, dp, dp, dp, C.c_double, C.c_double]
        L.orc_react_cvode.argtypes = [C.POINTER(self.CvodeParams), C.c_int, dp, dp, dp, dp, dp, dp, ip,
                                      C.c_double, C.c_double, C.POINTER(CvodeStats)]
        L.orc_fkernelspec.argtypes = [C.c_int, C.c_double, C.c_int, dp, dp, dp, dp, dp]
        mw = np.zeros(self.N)
        L.orc_mw(_p(mw))
        self.mw = mw

    # ---- unit probes -------------------------------------------------------------
    def pyt2re(self, P, Y, T):
        n = len(T)
        rho, e = np.zeros(n), np.zeros(n)
        self.lib.orc_pyt2re(n, _p(np.ascontiguousarray(P, dtype=float)), _p(np.ascontiguousarray(Y)), _p(np.ascontiguousarray(T, dtype=float)), _p(rho), _p(e))
        return rho, e

    def x2y(self, X):
        X = np.ascontiguousarray(np.atleast_2d(X), dtype=float)
        Y = np.zeros_like(X)
        self.lib.orc_x2y(X.shape[0], _p(X), _p(Y))
        return Y

    def ey2t(self, e, Y, Tguess):
        T = np.array(Tguess, dtype=float).copy()
        self.lib.orc_ey2t(len(T), _p(np.ascontiguousarray(e)), _p(np.ascontiguousarray(Y)), _p(T))
        return T

    def thermo(self, T, Y):
        n = len(T)
        ei, hi = np.zeros((n, self.N)), np.zeros((n, self.N))
        cv, em = np.zeros(n), np.zeros(n)
        self.lib.orc_thermo(n, _p(np.ascontiguousarray(T)), _p(np.ascontiguousarray(Y)), _p(ei), _p(cv), _p(em), _p(hi))
        return ei, cv, em, hi

    def ctoprim(self, u, small_num=1.0e-8):
        """pc_ctoprim: u [NVAR, n] -> q [QVAR, n], qa [6, n] (PeleC component layout, oracle/chem_oracle.c orc_ctoprim)"""
        u = np.ascontiguousarray(u, dtype=float)
        n = u.shape[1]
        q, qa = np.zeros((8 + self.N, n)), np.zeros((6, n))
        self.lib.orc_ctoprim.argtypes = [C.c_int, dp, dp, dp, C.c_double]
        self.lib.orc_ctoprim(n, _p(u), _p(q), _p(qa), float(small_num))
        return q, qa

    def transport(self, T, rho, Y):
        """SimpleTransport::transport: T [n], rho [n], Y [n, N] -> D [n, N], mu, xi, lam [n]"""
        T, rho, Y = (np.ascontiguousarray(x, dtype=float) for x in (T, rho, Y))
        n = len(T)
        D, mu, xi, lam = np.zeros((n, self.N)), np.zeros(n), np.zeros(n), np.zeros(n)
        self.lib.orc_transport.argtypes = [C.c_int, dp, dp, dp, dp, dp, dp, dp]
        self.lib.orc_transport(n, _p(T), _p(rho), _p(Y), _p(D), _p(mu), _p(xi), _p(lam))
        return D, mu, xi, lam

    def ckwc(self, T, Cc):
        w = np.zeros_like(Cc)
        self.lib.orc_ckwc(len(T), _p(np.ascontiguousarray(T)), _p(np.ascontiguousarray(Cc)), _p(w))
        return w

    def rty2wdot(self, rho, T, Y):
        w = np.zeros_like(Y)
        self.lib.orc_rty2wdot(len(T), _p(np.ascontiguousarray(rho)), _p(np.ascontiguousarray(T)), _p(np.ascontiguousarray(Y)), _p(w))
        return w

    def rty2jac(self, rho, T, Y, HP=0):
        n = len(T)
        J = np.zeros((n, self.neq * self.neq))
        self.lib.orc_rty2jac(n, _p(np.ascontiguousarray(rho)), _p(np.ascontiguousarray(T)), _p(np.ascontiguousarray(Y)), _p(J), HP)
        return J.reshape(n, self.neq, self.neq)  # [state, col, row]

    def fkernelspec(self, t, y, rhoe_init, rhoesrc, rYsrc, reactor_type=1):
        n = y.shape[0]
        ydot = np.zeros_like(y)
        self.lib.orc_fkernelspec(n, t, reactor_type, _p(np.ascontiguousarray(y)), _p(np.ascontiguousarray(rhoe_init)),
                                 _p(np.ascontiguousarray(rhoesrc)), _p(np.ascontiguousarray(rYsrc)), _p(ydot))
        return ydot

    # ---- reactors ----------------------------------------------------------------
    def react_rk64(self, st, dt, time=0.0, atol=1e-10, reactor_type=1, nsub=(10, 5, 500)):
        p = _RK64Params(reactor_type, atol, nsub[0], nsub[1], nsub[2])
        return self.lib.orc_react_rk64(C.byref(p), st.ncells, _p(st.rY), _p(st.rYsrc), _p(st.T), _p(st.rE),
                                       _p(st.rEsrc), _p(st.FC), dt, time)

    def react_cvode(self, st, dt, time=0.0, rtol=1e-10, atol=1e-10, max_order=2, solve_type="denseAJ_direct",
                    typ_vals=None, reactor_type=1, max_nls_iters=3, clean_init_massfrac=False, want_stats=False):
        p = self.CvodeParams()
        p.reactor_type, p.rtol, p.atol = reactor_type, rtol, atol
        for k in range(self.neq):
            p.typ_vals[k] = -1.0 if typ_vals is None else typ_vals[k]
        p.max_order, p.max_nls_iters = max_order, max_nls_iters
        p.analytic_jac = {"denseAJ_direct": 1, "dense_direct": 0}[solve_type]
        p.mxstep, p.maxnef, p.msbj = 10000, 100, 100
        p.clean_init_massfrac = int(clean_init_massfrac)
        stats = (CvodeStats * st.ncells)() if want_stats else None
        self.lib.orc_react_cvode(C.byref(p), st.ncells, _p(st.rY), _p(st.rYsrc), _p(st.T), _p(st.rE), _p(st.rEsrc),
                                 _p(st.FC), st.mask.ctypes.data_as(ip), dt, time, stats)
        return stats


class Reference:
    """The unmodified reference (PelePhysics reactors + CVODE) behind oracle/ref/peleref_capi.cpp."""
    kind = "reference"
    _initialised = set()

    def __init__(self, mech):
        path = os.path.join(HERE, "_ref", "libpeleref_%s.so" % mech)
        if not os.path.exists(path):
            raise FileNotFoundError(path + " (build with: make -C oracle/ref)")
        self.lib = L = C.CDLL(path, mode=C.RTLD_LOCAL)
        self.mech = mech
        # silence the reference's banner prints on stdout during init
        L.peleref_init()
        self.N = L.peleref_num_species()
        self.neq = self.N + 1
        self.nreac = L.peleref_num_reactions()
        buf = C.create_string_buffer(1 << 14)
        L.peleref_species_names(buf, len(buf))
        self.species = buf.value.decode().split()
        L.peleref_reactor_create.restype = C.c_void_p
        L.peleref_reactor_create.argtypes = [C.c_char_p, C.c_int, C.c_int]
        L.peleref_reactor_destroy.argtypes = [C.c_void_p]
        L.peleref_reactor_set_typ_vals.argtypes = [C.c_void_p, dp, C.c_int]
        L.peleref_react_box.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, dp, dp, dp, dp, dp, dp, ip, C.c_double, C.c_double]
        L.peleref_fkernelspec.argtypes = [C.c_int, C.c_double, C.c_int, dp, dp, dp, dp, dp]
        mw = np.zeros(self.N)
        L.peleref_mw(_p(mw))
        self.mw = mw

    def _parm(self, prefix, name, value):
        self.lib.peleref_set_parm(prefix.encode(), name.encode(), str(value).encode())

    def pyt2re(self, P, Y, T):
        n = len(T)
        rho, e = np.zeros(n), np.zeros(n)
        self.lib.peleref_pyt2re(n, _p(np.ascontiguousarray(P, dtype=float)), _p(np.ascontiguousarray(Y)), _p(np.ascontiguousarray(T, dtype=float)), _p(rho), _p(e))
        return rho, e

    def x2y(self, X):
        X = np.ascontiguousarray(np.atleast_2d(X), dtype=float)
        Y = np.zeros_like(X)
        self.lib.peleref_x2y(X.shape[0], _p(X), _p(Y))
        return Y

    def ey2t(self, e, Y, Tguess):
        T = np.array(Tguess, dtype=float).copy()
        self.lib.peleref_ey2t(len(T), _p(np.ascontiguousarray(e)), _p(np.ascontiguousarray(Y)), _p(T))
        return T

    def thermo(self, T, Y):
        n = len(T)
        ei, hi = np.zeros((n, self.N)), np.zeros((n, self.N))
        cv, em = np.zeros(n), np.zeros(n)
        self.lib.peleref_thermo(n, _p(np.ascontiguousarray(T)), _p(np.ascontiguousarray(Y)), _p(ei), _p(cv), _p(em), _p(hi))
        return ei, cv, em, hi

    def transport(self, T, rho, Y):
        """the reference's SimpleTransport::transport (peleref_transport)"""
        T, rho, Y = (np.ascontiguousarray(x, dtype=float) for x in (T, rho, Y))
        n = len(T)
        D, mu, xi, lam = np.zeros((n, self.N)), np.zeros(n), np.zeros(n), np.zeros(n)
        self.lib.peleref_transport.argtypes = [C.c_int, dp, dp, dp, dp, dp, dp, dp]
        self.lib.peleref_transport(n, _p(T), _p(rho), _p(Y), _p(D), _p(mu), _p(xi), _p(lam))
        return D, mu, xi, lam

    def ckwc(self, T, Cc):
        w = np.zeros_like(Cc)
        self.lib.peleref_ckwc(len(T), _p(np.ascontiguousarray(T)), _p(np.ascontiguousarray(Cc)), _p(w))
        return w

    def rty2wdot(self, rho, T, Y):
        w = np.zeros_like(Y)
        self.lib.peleref_rty2wdot(len(T), _p(np.ascontiguousarray(rho)), _p(np.ascontiguousarray(T)), _p(np.ascontiguousarray(Y)), _p(w))
        return w

    def rty2jac(self, rho, T, Y, HP=0):
        n = len(T)
        J = np.zeros((n, self.neq * self.neq))
        self.lib.peleref_rty2jac(n, _p(np.ascontiguousarray(rho)), _p(np.ascontiguousarray(T)), _p(np.ascontiguousarray(Y)), _p(J), HP)
        return J.reshape(n, self.neq, self.neq)

    def fkernelspec(self, t, y, rhoe_init, rhoesrc, rYsrc, reactor_type=1):
        n = y.shape[0]
        ydot = np.zeros_like(y)
        self.lib.peleref_fkernelspec(n, t, reactor_type, _p(np.ascontiguousarray(y)), _p(np.ascontiguousarray(rhoe_init)),
                                     _p(np.ascontiguousarray(rhoesrc)), _p(np.ascontiguousarray(rYsrc)), _p(ydot))
        return ydot

    def _react(self, key, st, dt, time, reactor_type, typ_vals=None):
        h = self.lib.peleref_reactor_create(key.encode(), reactor_type, 1)
        try:
            if typ_vals is not None:
                tv = np.ascontiguousarray(typ_vals, dtype=float)
                self.lib.peleref_reactor_set_typ_vals(h, _p(tv), len(tv))
            return self.lib.peleref_react_box(h, st.ncells, 1, 1, _p(st.rY), _p(st.rYsrc), _p(st.T), _p(st.rE),
                                              _p(st.rEsrc), _p(st.FC), st.mask.ctypes.data_as(ip), dt, time)
        finally:
            self.lib.peleref_reactor_destroy(h)

    def react_flat(self, key, y, ysrc, e, esrc, dt, time=0.0, reactor_type=1):
        """The reference's 1-D react(Real* rY_in, rYsrc_in, rX_in, rX_src_in, dt, time, ncells) (ReactorBase.H:42-54):
        y [ncells, N+1] (rhoY, T) in/out, ysrc [ncells, N], e [ncells] in/out, esrc [ncells].  Parameters (ode.*, cvode.*)
        as left by the last react_rk64 / react_cvode call or set through _parm.  Returns react()'s int."""
        ncells = y.shape[0]
        h = self.lib.peleref_reactor_create(key.encode(), reactor_type, ncells)
        try:
            self.lib.peleref_react_flat.restype = C.c_int
            self.lib.peleref_react_flat.argtypes = [C.c_void_p, dp, dp, dp, dp, C.c_double, C.c_double, C.c_int]
            return self.lib.peleref_react_flat(h, _p(y), _p(ysrc), _p(e), _p(esrc), dt, time, ncells)
        finally:
            self.lib.peleref_reactor_destroy(h)

    def react_rk64(self, st, dt, time=0.0, atol=1e-10, reactor_type=1, nsub=(10, 5, 500)):
        self._parm("ode", "atol", repr(atol))
        self._parm("ode", "rk64_nsubsteps_guess", nsub[0])
        self._parm("ode", "rk64_nsubsteps_min", nsub[1])
        self._parm("ode", "rk64_nsubsteps_max", nsub[2])
        self._parm("ode", "verbose", 0)
        return self._react("ReactorRK64", st, dt, time, reactor_type)

    def react_arkode(self, states, dt, time=0.0, rtol=1e-6, atol=1e-10, rk_method=40, rk_controller=0, typ_vals=None, reactor_type=1,
                     use_erkstep=1):
        """ReactorArkode (ERKStep) on a SEQUENCE of boxes with ONE reactor object, as PeleC's MFIter loop does: each state is
        one box (ncells x 1 x 1) integrated as one coupled system; the reactor's step-size controller is created once in init()
        and keeps its history from call to call (ReactorArkode.cpp:90-116, 207).  Returns react()'s ints (nfe per call)."""
        self._parm("ode", "rtol", repr(rtol))
        self._parm("ode", "atol", repr(atol))
        self._parm("ode", "verbose", 0)
        self._parm("ode", "use_erkstep", int(use_erkstep))
        self._parm("ode", "rk_method", rk_method)
        self._parm("ode", "rk_controller", rk_controller)
        self._parm("ode", "clean_init_massfrac", 0)
        h = self.lib.peleref_reactor_create(b"ReactorArkode", reactor_type, 1)
        out = []
        try:
            if typ_vals is not None:
                tv = np.ascontiguousarray(typ_vals, dtype=float)
                self.lib.peleref_reactor_set_typ_vals(h, _p(tv), len(tv))
            for st in states:
                out.append(self.lib.peleref_react_box(h, st.ncells, 1, 1, _p(st.rY), _p(st.rYsrc), _p(st.T), _p(st.rE),
                                                      _p(st.rEsrc), _p(st.FC), st.mask.ctypes.data_as(ip), dt, time))
        finally:
            self.lib.peleref_reactor_destroy(h)
        return out

    def react_cvode(self, st, dt, time=0.0, rtol=1e-10, atol=1e-10, max_order=2, solve_type="denseAJ_direct",
                    typ_vals=None, reactor_type=1, max_nls_iters=3, clean_init_massfrac=False, want_stats=False):
        self._parm("ode", "rtol", repr(rtol))
        self._parm("ode", "atol", repr(atol))
        self._parm("ode", "verbose", 0)
        self._parm("ode", "max_nls_iters", max_nls_iters)
        self._parm("ode", "clean_init_massfrac", int(clean_init_massfrac))
        self._parm("cvode", "solve_type", solve_type)
        self._parm("cvode", "max_order", max_order)
        self._react("ReactorCvode", st, dt, time, reactor_type, typ_vals)
        return None


# ----------------------------------------------------------------------------------------------
# Synthetic states
# ----------------------------------------------------------------------------------------------
def reacteval_states(lib, ncells_xyz, fuel=None):
    """Set A: ReactEval's initialize_data (PP/Testing/Exec/ReactEval/GPU_misc.H:13-82) on the domain
    [-1,1]^3 (inputs.3d_Array4 geometry): T(y) = 2000 + 500*y/L + 100*sin(2*pi*y/(L/4)), P = 1 atm,
    X_fuel = 0.1, X_O2 = 0.2, X_N2 = 0.7; zero external sources; energy type 1 (rho*e)."""
    nx, ny, nz = ncells_xyz
    fuel = fuel or FUEL.get(lib.mech) or next(f for f in ("NC12H26", "NC7H16", "C2H4", "CH4", "H2") if f in lib.species)
    X = np.zeros(lib.N)
    X[lib.species.index("O2")] = 0.2
    X[lib.species.index(fuel)] = 0.1
    X[lib.species.index("N2")] = 0.7
    Y = lib.x2y(X)[0]
    plo, phi = -1.0, 1.0
    Ly = phi - plo
    dy = Ly / ny
    j = np.arange(ny)
    y = plo + (j + 0.5) * dy
    Tj = 2000.0 + (2500.0 - 2000.0) * y / Ly + 100.0 * np.sin(2.0 * 3.1415926535897932 * y / (Ly / 4.0))
    rho_j, e_j = lib.pyt2re(np.full(ny, 1013250.0), np.tile(Y, (ny, 1)), Tj)
    # cell (i,j,k) -> flat index (k*ny + j)*nx + i
    jj = np.tile(np.repeat(np.arange(ny), nx), nz)
    rY = (Y[:, None] * rho_j[jj][None, :])
    return BoxState(rY, Tj[jj], (rho_j * e_j)[jj])


# ----------------------------------------------------------------------------------------------
# PeleC::react_state body (test infrastructure; restated from PC/Source/React.cpp:181-343, numpy)
# PINNED against PeleC::react_state of the unmodified application (oracle/ref_pelec, tests/golden/make_golden3.py,
# tests/test_react_state_golden.py): bit for bit with ReactorRK64, incl. the rotating-frame terms (pelec.do_rf).
# ----------------------------------------------------------------------------------------------
UIDX = dict(urho=0, umx=1, umy=2, umz=3, ueden=4, ueint=5, utemp=6, ufs=7)


def rot_energy(idx, omega, axis, axis_loc, prob_lo, dx):
    """get_rot_energy (PC/Source/Utilities.H:23-65): idx [3, ncells] global cell indices -> 0.5 omega^2 r^2 per cell"""
    r = [(prob_lo[d] + (idx[d] + 0.5) * dx[d]) - axis_loc[d] for d in range(3)]
    rad2 = r[0] * r[0] + r[1] * r[1] + r[2] * r[2]
    rad2 = rad2 - r[axis] * r[axis]
    rad = np.sqrt(rad2)
    return 0.5 * omega * omega * rad * rad


def react_state_oracle(orc, sold, snew, nonrs, dt, do_update=True, reactor="cvode", rotenrg=0.0, **kw):
    """sold, snew, nonrs: [NVAR, ncells] arrays (NVAR = 7 + N).  Returns (snew_out, I_R [N+2, ncells], FC).
    rotenrg: rotating-frame energy per cell (rot_energy), 0 without pelec.do_rf."""
    u = UIDX
    N = orc.N
    sold, snew, nonrs = (np.array(x, dtype=np.float64) for x in (sold, snew, nonrs))
    nc = sold.shape[1]
    # ---- React.cpp:181-220 (pack)
    rhou, rhov, rhow, rho_old = sold[u["umx"]], sold[u["umy"]], sold[u["umz"]], sold[u["urho"]]
    rhoInv = 1.0 / rho_old
    e_old = (sold[u["ueden"]] - 0.5 * (rhou * rhou + rhov * rhov + rhow * rhow) * rhoInv) * rhoInv
    e_old = e_old + rotenrg
    rhou, rhov, rhow = snew[u["umx"]], snew[u["umy"]], snew[u["umz"]]
    rhoInv = 1.0 / snew[u["urho"]]
    rhoedot_ext = (snew[u["ueden"]] - 0.5 * (rhou * rhou + rhov * rhov + rhow * rhow) * rhoInv + snew[u["urho"]] * rotenrg
                   - rho_old * e_old) / dt
    st = BoxState(sold[u["ufs"]:u["ufs"] + N].copy(), sold[u["utemp"]].copy(), sold[u["ueint"]].copy(),
                  nonrs[u["ufs"]:u["ufs"] + N].copy(), rhoedot_ext.copy())
    # ---- React.cpp:222-228 (reactor, current_time = 0)
    if reactor == "cvode":
        orc.react_cvode(st, dt, 0.0, **kw)
    else:
        orc.react_rk64(st, dt, 0.0, **kw)
    # ---- React.cpp:233-321 (unpack, I_R)
    umnew = sold[u["umx"]] + dt * nonrs[u["umx"]]
    vmnew = sold[u["umy"]] + dt * nonrs[u["umy"]]
    wmnew = sold[u["umz"]] + dt * nonrs[u["umz"]]
    rhonew = np.zeros(nc)
    for n in range(N):
        rhonew = rhonew + st.rY[n]
    out = snew.copy()
    eint_new = rho_old * e_old + dt * rhoedot_ext
    eden_new = eint_new + 0.5 * (umnew * umnew + vmnew * vmnew + wmnew * wmnew) / rhonew - rhonew * rotenrg
    if do_update:
        out[u["urho"]], out[u["umx"]], out[u["umy"]], out[u["umz"]] = rhonew, umnew, vmnew, wmnew
        out[u["ufs"]:u["ufs"] + N] = st.rY
        out[u["utemp"]] = st.T
        out[u["ueint"]] = eint_new
        out[u["ueden"]] = eden_new
    I_R = np.zeros((N + 2, nc))
    for n in range(N):
        I_R[n] = (st.rY[n] - sold[u["ufs"] + n]) / dt - nonrs[u["ufs"] + n]
    I_R[N] = (eden_new - sold[u["ueden"]]) / dt - nonrs[u["ueden"]]
    # ---- React.cpp:325-343 (heat release from the NEW state: RTY2Hi depends on T only)
    Th = out[u["utemp"]]
    Y = (out[u["ufs"]:u["ufs"] + N] / out[u["urho"]]).T.copy()
    hi = orc.thermo(Th, Y)[3]  # [ncells, N] species enthalpies, mass units (CKHMS)
    for n in range(N):
        I_R[N + 1] -= hi[:, n] * I_R[n]
    return out, I_R, st.FC
