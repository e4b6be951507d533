"""Host-side mirror of the reference's reactor interface over the C ABI (include/pelerx_b200.h).

``ReactorB200`` keeps the names, argument meaning and error behaviour of
``pele::physics::reactions::ReactorBase`` (PP/Source/Reactions/ReactorBase.H:14-106):

    r = ReactorB200.create("ReactorCvode", mechanism="LiDryer", **parm)   # ReactorBase::create(key)
    r.init(reactor_type=1, ncells=1)                                      # parses ode.* / cvode.* keys
    r.set_typ_vals_ode(typ)                                               # optional
    r.react(box, rY, rYsrc, T, rEner, rEner_src, FC, mask, dt, time)      # Array4 version
    r.close()

PyTorch is only used for device memory and streams; every number is computed by the CUDA
kernels in ``pelec_b200/lib/libpelerx_b200.so``.  There is no CPU fallback: if the library or a
CUDA device is missing, construction raises.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PRX_LIB_PATH") or os.path.join(HERE, "lib", "libpelerx_b200.so")  # override: A/B builds

MECHANISMS = {"LiDryer": 0, "drm19": 1, "dodecane_lu": 2}
REACTOR_KINDS = {"ReactorRK64": 0, "ReactorCvode": 1, "ReactorArkode": 2}
SOLVE_TYPES = {"denseAJ_direct": 0, "dense_direct": 1}
# solve types of the reference that this path does not implement (PP ReactorCvode.cpp:507-934)
UNSUPPORTED_SOLVE_TYPES = ("GMRES", "precGMRES", "sparse_direct", "custom_direct", "magma_direct", "fixed_point")

PRX_OK = 0
STATUS = {0: "PRX_OK", -1: "PRX_ERR_BAD_ARG", -2: "PRX_ERR_NO_DEVICE", -3: "PRX_ERR_CUDA",
          -4: "PRX_ERR_UNSUPPORTED", -5: "PRX_ERR_ALLOC"}


class PrxError(RuntimeError):
    def __init__(self, code, msg=""):
        self.code = code
        super().__init__("%s%s" % (STATUS.get(code, str(code)), (": " + msg) if msg else ""))


class prx_config(C.Structure):
    _fields_ = [("abi_version", C.c_int), ("mechanism", C.c_int), ("reactor_kind", C.c_int),
                ("reactor_type", C.c_int), ("rtol", C.c_double), ("atol", C.c_double),
                ("max_order", C.c_int), ("max_nls_iters", C.c_int), ("solve_type", C.c_int),
                ("rk64_nsubsteps_guess", C.c_int), ("rk64_nsubsteps_min", C.c_int),
                ("rk64_nsubsteps_max", C.c_int), ("clean_init_massfrac", C.c_int), ("verbose", C.c_int),
                ("device", C.c_int), ("ark_method", C.c_int), ("ark_controller", C.c_int), ("ark_use_erkstep", C.c_int)]


class prx_array4(C.Structure):
    _fields_ = [("p", C.c_void_p), ("jstride", C.c_longlong), ("kstride", C.c_longlong),
                ("nstride", C.c_longlong), ("begin", C.c_int * 3), ("end", C.c_int * 3), ("ncomp", C.c_int)]


class prx_state_index(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("urho", "umx", "umy", "umz", "ueden", "ueint", "utemp", "ufs", "nvar")]


class prx_stats(C.Structure):
    _fields_ = [(n, C.c_longlong) for n in
                ("cells", "cells_masked", "n_steps", "n_rhs", "n_jac", "n_setups", "n_newton", "n_errfail",
                 "n_convfail", "n_failed", "max_steps", "reserved")]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_ if n != "reserved"}


class prx_prim_index(C.Structure):
    """PC/Source/IndexDefines.H: primitive (QRHO ... QFS, QVAR) and auxiliary (QGAMC ... QRSPEC, NQAUX) component indices"""
    _fields_ = [(n, C.c_int) for n in ("qrho", "qu", "qv", "qw", "qgame", "qpres", "qreint", "qtemp", "qfs", "qvar",
                                       "qgamc", "qc", "qcsml", "qdpdr", "qdpde", "qrspec", "nqaux")]


_lib = None


def load_library():
    """dlopen the in-tree CUDA library; fails loudly if it was not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FileNotFoundError(
            LIB_PATH + " is missing: build it with __graft_entry__.build() (nvcc, sm_100a). "
            "There is no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp, dp, ip = C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int)
    a4 = C.POINTER(prx_array4)
    L.prx_create.argtypes = [C.POINTER(prx_config), C.POINTER(vp)]
    L.prx_destroy.argtypes = [vp]
    L.prx_set_typ_vals.argtypes = [vp, dp, C.c_int]
    L.prx_last_error.argtypes = [vp]
    L.prx_last_error.restype = C.c_char_p
    L.prx_mechanism_name.restype = C.c_char_p
    L.prx_default_config.argtypes = [C.POINTER(prx_config)]
    L.prx_react_box.argtypes = [vp, ip, ip, a4, a4, a4, a4, a4, a4, a4, C.c_double, C.c_double, vp]
    L.prx_react_flat.argtypes = [vp, vp, vp, vp, vp, vp, C.c_double, C.c_double, C.c_int, vp]
    L.prx_react_boxes.argtypes = [vp, C.c_int, ip, ip, a4, a4, a4, a4, a4, a4, a4, C.c_double, C.c_double, vp]
    sx = C.POINTER(prx_state_index)
    L.prx_react_state_box.argtypes = [vp, ip, ip, a4, a4, a4, a4, a4, sx, C.c_double, C.c_int, vp]
    L.prx_react_state_boxes.argtypes = [vp, C.c_int, ip, ip, a4, a4, a4, a4, a4, sx, C.c_double, C.c_int, vp]
    L.prx_compute_temp_boxes.argtypes = [vp, C.c_int, ip, ip, a4, sx, vp]
    L.prx_ctoprim_boxes.argtypes = [vp, C.c_int, ip, ip, a4, a4, a4, sx, C.POINTER(prx_prim_index), C.c_double, vp]
    L.prx_transport_boxes.argtypes = [vp, C.c_int, ip, ip, a4, a4, a4, a4, a4, a4, a4, vp]
    L.prx_flatten.argtypes = [vp, ip, ip, C.c_int, C.c_int, a4, a4, a4, a4, a4, vp, vp, vp, vp, vp]
    L.prx_unflatten.argtypes = [vp, ip, ip, C.c_int, C.c_int, a4, a4, a4, a4, a4, vp, vp, vp, C.c_double, vp]
    L.prx_react_flat_host.argtypes = [vp, vp, vp, vp, vp, vp, C.c_double, C.c_double, C.c_int]
    L.prx_find_mechanism.argtypes = [C.c_int, C.c_int, C.c_char_p, C.c_char_p]
    L.prx_species_name.argtypes = [C.c_int, C.c_int]
    L.prx_species_name.restype = C.c_char_p
    L.prx_react_box_host.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp, vp, vp, vp, vp, vp, vp, C.c_double, C.c_double]
    L.prx_reset_stats.argtypes = [vp]
    L.prx_get_stats.argtypes = [vp, C.POINTER(prx_stats)]
    L.prx_last_react_return.argtypes = [vp]
    L.prx_kernel_name.argtypes = [vp]
    L.prx_kernel_name.restype = C.c_char_p
    L.prx_probe_rhs_host.argtypes = [vp, C.c_int, C.c_double, vp, vp, vp, vp, vp]
    L.prx_probe_jac_host.argtypes = [vp, C.c_int, vp, vp]
    _lib = L
    return L


class Box:
    """Inclusive index box, the subset of amrex::Box the reactor needs."""

    def __init__(self, lo, hi):
        self.lo = tuple(int(v) for v in lo)
        self.hi = tuple(int(v) for v in hi)

    @property
    def length(self):
        return tuple(max(0, h - l + 1) for l, h in zip(self.lo, self.hi))

    def numPts(self):
        n = 1
        for v in self.length:
            n *= v
        return n


class Array4:
    """amrex::Array4<T> over a torch CUDA tensor of shape [ncomp, nz, ny, nx] (i fastest), with
    `begin` = index of element [.,0,0,0] (AMR/Src/Base/AMReX_Array4.H:60-145)."""

    def __init__(self, tensor, begin=(0, 0, 0)):
        if tensor.dim() == 3:
            tensor = tensor.unsqueeze(0)
        assert tensor.dim() == 4 and tensor.stride(3) == 1, "Array4 needs a [ncomp, nz, ny, nx] tensor with unit stride in i"
        self.t = tensor
        self.begin = tuple(int(b) for b in begin)

    def c_struct(self):
        t = self.t
        nc, nz, ny, nx = t.shape
        es = t.element_size()
        a = prx_array4()
        a.p = t.data_ptr()
        assert t.stride(3) == 1, "i must be the unit-stride index"
        a.jstride, a.kstride, a.nstride = t.stride(2), t.stride(1), t.stride(0)
        for d in range(3):
            a.begin[d] = self.begin[d]
        a.end[0], a.end[1], a.end[2] = self.begin[0] + nx, self.begin[1] + ny, self.begin[2] + nz
        a.ncomp = nc
        del es
        return a


class ReactorB200:
    """B200 reactor behind the ReactorBase interface."""

    @staticmethod
    def identifier():
        return "ReactorB200"

    @classmethod
    def create(cls, key, mechanism="LiDryer", **parm):
        """ReactorBase::create(key) (PP/Source/Factory.H:72-79): unknown keys raise, listing valid ones."""
        if key not in REACTOR_KINDS:
            raise KeyError("Unknown reactor '%s'. Valid: %s" % (key, ", ".join(sorted(REACTOR_KINDS))))
        return cls(key, mechanism, **parm)

    def __init__(self, key, mechanism, **parm):
        if mechanism not in MECHANISMS:  # a mechanism beyond the three ABI constants: ask the library of this build
            lib = load_library()
            lib.prx_mechanism_id.argtypes = [C.c_char_p]
            mid = lib.prx_mechanism_id(mechanism.encode())
            if mid < 0:
                valid = [lib.prx_mechanism_name(i).decode() for i in range(lib.prx_num_mechanisms())]
                raise KeyError("Unknown mechanism '%s'. This build holds: %s" % (mechanism, ", ".join(valid)))
            MECHANISMS[mechanism] = mid
        self.key = key
        self.mechanism = mechanism
        # ParmParse-style keys with the reference's defaults (ReactorCvode.cpp:18-35, ReactorRK64.cpp:12-18)
        self.parm = {"ode.rtol": 1e-6, "ode.atol": 1e-10, "ode.max_nls_iters": 3, "ode.clean_init_massfrac": 0,
                     "ode.verbose": 0, "cvode.max_order": 2, "cvode.solve_type": "denseAJ_direct",
                     "ode.rk64_nsubsteps_guess": 10, "ode.rk64_nsubsteps_min": 5, "ode.rk64_nsubsteps_max": 500,
                     "ode.rk_method": 40, "ode.rk_controller": 0, "ode.use_erkstep": 1}
        for k, v in parm.items():
            k = k.replace("__", ".")
            if k not in self.parm:
                raise KeyError("unknown parameter " + k)
            self.parm[k] = v
        self.h = None
        self.lib = load_library()
        self.N = self.lib.prx_num_species(MECHANISMS[mechanism])

    def init(self, reactor_type, ncells=1, device=-1):
        st = self.parm["cvode.solve_type"]
        if st in UNSUPPORTED_SOLVE_TYPES:
            raise PrxError(-4, "cvode.solve_type=%s is not implemented by the B200 path" % st)
        if st not in SOLVE_TYPES:
            raise PrxError(-1, "Wrong solve_type '%s'" % st)
        cfg = prx_config()
        self.lib.prx_default_config(C.byref(cfg))
        cfg.mechanism = MECHANISMS[self.mechanism]
        cfg.reactor_kind = REACTOR_KINDS[self.key]
        cfg.reactor_type = int(reactor_type)
        cfg.rtol, cfg.atol = float(self.parm["ode.rtol"]), float(self.parm["ode.atol"])
        cfg.max_order = int(self.parm["cvode.max_order"])
        cfg.max_nls_iters = int(self.parm["ode.max_nls_iters"])
        cfg.solve_type = SOLVE_TYPES[st]
        cfg.rk64_nsubsteps_guess = int(self.parm["ode.rk64_nsubsteps_guess"])
        cfg.rk64_nsubsteps_min = int(self.parm["ode.rk64_nsubsteps_min"])
        cfg.rk64_nsubsteps_max = int(self.parm["ode.rk64_nsubsteps_max"])
        cfg.clean_init_massfrac = int(self.parm["ode.clean_init_massfrac"])
        cfg.verbose = int(self.parm["ode.verbose"])
        cfg.device = int(device)
        cfg.ark_method, cfg.ark_controller = int(self.parm["ode.rk_method"]), int(self.parm["ode.rk_controller"])
        cfg.ark_use_erkstep = int(self.parm["ode.use_erkstep"])
        h = C.c_void_p()
        rc = self.lib.prx_create(C.byref(cfg), C.byref(h))
        if rc != PRX_OK:
            raise PrxError(rc, "prx_create")
        self.h = h
        self.cfg = cfg
        return 0

    def _check(self, rc):
        if rc != PRX_OK:
            raise PrxError(rc, (self.lib.prx_last_error(self.h) or b"").decode())

    def close(self):
        if self.h is not None:
            self.lib.prx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_typ_vals_ode(self, typ_vals):
        if len(typ_vals) != self.N + 1:
            raise PrxError(-1, "typ_vals must have NUM_SPECIES+1 entries")
        arr = (C.c_double * (self.N + 1))(*[float(v) for v in typ_vals])
        self._check(self.lib.prx_set_typ_vals(self.h, arr, self.N + 1))

    # -- Array4 react on device tensors ------------------------------------------------------
    def react(self, box, rY_in, rYsrc_in, T_in, rEner_in, rEner_src_in, FC_in, mask, dt_react, time, stream=None):
        lo = (C.c_int * 3)(*box.lo)
        hi = (C.c_int * 3)(*box.hi)
        structs = [a.c_struct() for a in (rY_in, rYsrc_in, T_in, rEner_in, rEner_src_in, FC_in)]
        m = mask.c_struct() if mask is not None else None
        sp = C.c_void_p(stream) if stream else None
        self._check(self.lib.prx_react_box(
            self.h, lo, hi, *[C.byref(s) for s in structs], C.byref(m) if m is not None else None,
            float(dt_react), float(time), sp))
        return 0

    # -- several boxes in one call (the caller's MFIter loop) ------------------------------------
    def react_boxes(self, boxes, rY, rYsrc, T, rE, rEsrc, FC, masks, dt_react, time, stream=None):
        """boxes: list of Box; every array argument: list of Array4, one per box (masks may be None)."""
        nb = len(boxes)
        lo = (C.c_int * (3 * nb))(*[v for b in boxes for v in b.lo])
        hi = (C.c_int * (3 * nb))(*[v for b in boxes for v in b.hi])
        arrs = [(prx_array4 * nb)(*[a.c_struct() for a in lst]) for lst in (rY, rYsrc, T, rE, rEsrc, FC)]
        m = (prx_array4 * nb)(*[a.c_struct() for a in masks]) if masks is not None else None
        sp = C.c_void_p(stream) if stream else None
        self._check(self.lib.prx_react_boxes(self.h, nb, lo, hi, *arrs, m, float(dt_react), float(time), sp))
        return 0

    # -- fused PeleC::react_state body (PC/Source/React.cpp:140-343) ----------------------------
    @staticmethod
    def state_index(nspec, urho=0, umx=1, umy=2, umz=3, ueden=4, ueint=5, utemp=6, ufs=7, nvar=None):
        """PC/Source/IndexDefines.H:3-12 (NUM_ADV = NUM_AUX = 0: UFS = 7, NVAR = 7 + NUM_SPECIES)"""
        ix = prx_state_index()
        ix.urho, ix.umx, ix.umy, ix.umz, ix.ueden, ix.ueint, ix.utemp, ix.ufs = urho, umx, umy, umz, ueden, ueint, utemp, ufs
        ix.nvar = ufs + nspec if nvar is None else nvar
        return ix

    def set_rotating_frame(self, omega, axis, axis_loc, prob_lo, dx, on=True):
        """pelec.do_rf / rf_omega / rf_axis / rf_axis_{x,y,z} + the level's geometry (PC/Source/React.cpp:123-131)"""
        d3 = C.c_double * 3
        self.lib.prx_set_rotating_frame.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_int, d3, d3, d3]
        self._check(self.lib.prx_set_rotating_frame(self.h, int(bool(on)), float(omega), int(axis), d3(*axis_loc), d3(*prob_lo), d3(*dx)))

    def react_state(self, box, sold, snew, nonrs, I_R, dt, do_update=True, work=None, ix=None, stream=None):
        lo = (C.c_int * 3)(*box.lo)
        hi = (C.c_int * 3)(*box.hi)
        structs = [a.c_struct() for a in (sold, snew, nonrs, I_R)]
        w = work.c_struct() if work is not None else None
        ix = ix or self.state_index(self.N)
        sp = C.c_void_p(stream) if stream else None
        self._check(self.lib.prx_react_state_box(
            self.h, lo, hi, *[C.byref(x) for x in structs], C.byref(w) if w is not None else None, C.byref(ix),
            float(dt), int(bool(do_update)), sp))
        return 0

    def react_state_boxes(self, boxes, sold, snew, nonrs, I_R, dt, do_update=True, work=None, ix=None, stream=None):
        nb = len(boxes)
        lo = (C.c_int * (3 * nb))(*[v for b in boxes for v in b.lo])
        hi = (C.c_int * (3 * nb))(*[v for b in boxes for v in b.hi])
        arrs = [(prx_array4 * nb)(*[a.c_struct() for a in lst]) for lst in (sold, snew, nonrs, I_R)]
        w = (prx_array4 * nb)(*[a.c_struct() for a in work]) if work is not None else None
        ix = ix or self.state_index(self.N)
        sp = C.c_void_p(stream) if stream else None
        self._check(self.lib.prx_react_state_boxes(self.h, nb, lo, hi, *arrs, w, C.byref(ix), float(dt),
                                                   int(bool(do_update)), sp))
        return 0

    # -- PeleC::computeTemp's kernel (pc_cmpTemp) -------------------------------------------------
    def compute_temp(self, boxes, states, ix=None, stream=None):
        nb = len(boxes)
        lo = (C.c_int * (3 * nb))(*[v for b in boxes for v in b.lo])
        hi = (C.c_int * (3 * nb))(*[v for b in boxes for v in b.hi])
        arr = (prx_array4 * nb)(*[a.c_struct() for a in states])
        ix = ix or self.state_index(self.N)
        sp = C.c_void_p(stream) if stream else None
        self._check(self.lib.prx_compute_temp_boxes(self.h, nb, lo, hi, arr, C.byref(ix), sp))
        return 0

    # -- pc_ctoprim and SimpleTransport::get_transport_coeffs ---------------------------
    @staticmethod
    def prim_index(nspec):
        """PeleC's layout with NUM_ADV = NUM_AUX = NUM_LIN = 0 (PC/Source/IndexDefines.H:15-30,71-78)"""
        qx = prx_prim_index()
        (qx.qrho, qx.qu, qx.qv, qx.qw, qx.qgame, qx.qpres, qx.qreint, qx.qtemp, qx.qfs) = range(9)
        qx.qvar = 8 + nspec
        (qx.qgamc, qx.qc, qx.qcsml, qx.qdpdr, qx.qdpde, qx.qrspec) = range(6)
        qx.nqaux = 6
        return qx

    @staticmethod
    def _boxes(boxes):
        nb = len(boxes)
        return nb, (C.c_int * (3 * nb))(*[v for b in boxes for v in b.lo]), (C.c_int * (3 * nb))(*[v for b in boxes for v in b.hi])

    def ctoprim(self, boxes, u, q, qaux, ix=None, qx=None, small_num=1.0e-8, stream=None):
        nb, lo, hi = self._boxes(boxes)
        arrs = [(prx_array4 * nb)(*[a.c_struct() for a in lst]) for lst in (u, q, qaux)]
        ix = ix or self.state_index(self.N)
        qx = qx or self.prim_index(self.N)
        sp = C.c_void_p(stream) if stream else None
        self._check(self.lib.prx_ctoprim_boxes(self.h, nb, lo, hi, *arrs, C.byref(ix), C.byref(qx), float(small_num), sp))
        return 0

    def transport(self, boxes, Y, T, rho, D, mu, xi, lam, stream=None):
        nb, lo, hi = self._boxes(boxes)
        arrs = [(prx_array4 * nb)(*[a.c_struct() for a in lst]) for lst in (Y, T, rho, D, mu, xi, lam)]
        sp = C.c_void_p(stream) if stream else None
        self._check(self.lib.prx_transport_boxes(self.h, nb, lo, hi, *arrs, sp))
        return 0

    # -- flat react on device tensors ----------------------------------------------------------
    def react_flat(self, y, ysrc, e, esrc, dt_react, time, ncells, fc=None, stream=None):
        sp = C.c_void_p(stream) if stream else None
        self._check(self.lib.prx_react_flat(
            self.h, y.data_ptr(), ysrc.data_ptr(), e.data_ptr(), esrc.data_ptr(),
            fc.data_ptr() if fc is not None else None, float(dt_react), float(time), int(ncells), sp))
        return 0

    # -- ReactorBase::flatten / unflatten on device tensors (ordering 0 = YCOrder, 1 = CYOrder) --------
    def flatten(self, box, ncells, rhoY, frcExt, temperature, rhoE, frcEExt, y_vect, src_vect, vect_energy,
                src_vect_energy, ordering=0, stream=None):
        lo, hi = (C.c_int * 3)(*box.lo), (C.c_int * 3)(*box.hi)
        st = [a.c_struct() for a in (rhoY, frcExt, temperature, rhoE, frcEExt)]
        self._check(self.lib.prx_flatten(
            self.h, lo, hi, int(ncells), int(ordering), *[C.byref(x) for x in st], y_vect.data_ptr(), src_vect.data_ptr(),
            vect_energy.data_ptr(), src_vect_energy.data_ptr(), C.c_void_p(stream) if stream else None))

    def unflatten(self, box, ncells, rhoY, temperature, rhoE, frcEExt, FC_in, y_vect, vect_energy, FCunt, dt,
                  ordering=0, stream=None):
        lo, hi = (C.c_int * 3)(*box.lo), (C.c_int * 3)(*box.hi)
        st = [a.c_struct() for a in (rhoY, temperature, rhoE, frcEExt, FC_in)]
        self._check(self.lib.prx_unflatten(
            self.h, lo, hi, int(ncells), int(ordering), *[C.byref(x) for x in st], y_vect.data_ptr(), vect_energy.data_ptr(),
            FCunt.data_ptr(), float(dt), C.c_void_p(stream) if stream else None))

    def react_flat_host(self, y, ysrc, e, esrc, dt_react, time, ncells, fc=None):
        """1-D react with HOST numpy arrays (the reference's flat react takes host pointers)."""
        self._check(self.lib.prx_react_flat_host(
            self.h, y.ctypes.data, ysrc.ctypes.data, e.ctypes.data, esrc.ctypes.data,
            fc.ctypes.data if fc is not None else None, float(dt_react), float(time), int(ncells)))
        return 0

    # -- host (numpy / pinned torch) convenience: H2D + react + D2H ---------------------------------
    def react_host(self, nx, ny, nz, rY, rYsrc, T, rE, rEsrc, FC, mask, dt_react, time):
        def ptr(a):
            if a is None:
                return None
            if hasattr(a, "data_ptr"):
                return a.data_ptr()
            return a.ctypes.data
        self._check(self.lib.prx_react_box_host(
            self.h, nx, ny, nz, ptr(rY), ptr(rYsrc), ptr(T), ptr(rE), ptr(rEsrc), ptr(FC), ptr(mask),
            float(dt_react), float(time)))
        return 0

    def stats(self):
        s = prx_stats()
        self._check(self.lib.prx_get_stats(self.h, C.byref(s)))
        return s.as_dict()

    def reset_stats(self):
        self._check(self.lib.prx_reset_stats(self.h))

    def kernel_name(self):
        return self.lib.prx_kernel_name(self.h).decode()

    def last_react_return(self):
        return self.lib.prx_last_react_return(self.h)

    def probe_rhs(self, t, y, rhoe, rhoesrc, rYsrc):
        import numpy as np
        # contiguous float64 copies held in locals until the call returns (a temporary's .ctypes.data would dangle)
        y = np.ascontiguousarray(y, dtype=np.float64)
        rhoe = np.ascontiguousarray(rhoe, dtype=np.float64)
        rhoesrc = np.ascontiguousarray(rhoesrc, dtype=np.float64)
        rYsrc = np.ascontiguousarray(rYsrc, dtype=np.float64)
        n = y.shape[0]
        if rhoe.size != n or rhoesrc.size != n or rYsrc.size != n * self.N or y.shape[1] != self.N + 1:
            raise PrxError(-1, "probe_rhs: array shapes do not match (n, NUM_SPECIES+1)")
        out = np.zeros_like(y)
        self._check(self.lib.prx_probe_rhs_host(
            self.h, n, float(t), y.ctypes.data, rhoe.ctypes.data, rhoesrc.ctypes.data, rYsrc.ctypes.data, out.ctypes.data))
        return out

    def probe_jac(self, y):
        import numpy as np
        y = np.ascontiguousarray(y, dtype=np.float64)
        n, neq = y.shape
        J = np.zeros((n, neq, neq))
        self._check(self.lib.prx_probe_jac_host(self.h, n, y.ctypes.data, J.ctypes.data))
        return J  # [state, col, row]
