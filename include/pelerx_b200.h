/* pelerx_b200.h - C ABI of the B200-native (sm_100a) stiff-chemistry reactor.
 *
 * This is the drop-in boundary for ONE hot path of PeleC / PelePhysics: the pointwise
 * chemistry advance behind pele::physics::reactions::ReactorBase
 * (PP/Source/Reactions/ReactorBase.H:14-106), i.e. what PeleC::react_state
 * (PC/Source/React.cpp:222-228) and ReactEval (PP/Testing/Exec/ReactEval/main.cpp:516-522)
 * call once per AMReX box.  A thin C++ shim (integration/ReactorB200.{H,cpp}) derives from
 * ReactorBase::Register<> and only unpacks amrex::Box / amrex::Array4 into the PODs below.
 *
 * Conventions
 *   - every entry point returns 0 on success, a negative prx_status otherwise; nothing
 *     aborts (the reference calls amrex::Abort, PP/Source/Reactions/ReactorUtils.cpp:12-46);
 *   - all array pointers are DEVICE pointers unless the function name ends in _host;
 *   - calls are asynchronous on the given CUDA stream (the caller synchronises, as
 *     PC/Source/React.cpp:230 does) except prx_get_stats / *_host which synchronise;
 *   - there is NO CPU fallback: without a CUDA device prx_create fails with PRX_ERR_NO_DEVICE.
 *
 * (PC = /root/reference, PP = PC/Submodules/PelePhysics.)
 */
#ifndef PELERX_B200_H
#define PELERX_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define PRX_ABI_VERSION 2

#if defined(__GNUC__)
#define PRX_API __attribute__((visibility("default")))
#else
#define PRX_API
#endif

typedef enum {
  PRX_OK = 0,
  PRX_ERR_BAD_ARG = -1,      /* null pointer, bad size, unknown enum value */
  PRX_ERR_NO_DEVICE = -2,    /* no CUDA device / driver: the library never computes on the CPU */
  PRX_ERR_CUDA = -3,         /* a CUDA runtime call failed; see prx_last_error */
  PRX_ERR_UNSUPPORTED = -4,  /* option the reference has but this path does not implement */
  PRX_ERR_ALLOC = -5
} prx_status;

/* compile-time mechanisms (the reference selects one per executable through
 * Chemistry_Model / NUM_SPECIES, PC/Exec/Make.PeleC:75-81; here all are linked in) */
typedef enum {
  PRX_MECH_LIDRYER = 0,      /* PP/Mechanisms/LiDryer      9 species, 21 reactions  */
  PRX_MECH_DRM19 = 1,        /* PP/Mechanisms/drm19       21 species, 84 reactions  */
  PRX_MECH_DODECANE_LU = 2   /* PP/Mechanisms/dodecane_lu 53 species, 268 reactions */
} prx_mechanism;

/* pelec.chem_integrator / chem_integrator keys (PC/Source/Params/_cpp_parameters:302-305) */
typedef enum {
  PRX_REACTOR_RK64 = 0,      /* "ReactorRK64"  PP/Source/Reactions/ReactorRK64.cpp:166-320  */
  PRX_REACTOR_CVODE = 1,     /* "ReactorCvode" PP/Source/Reactions/ReactorCvode.cpp:1267-1489,
                                per-cell (CPU-branch) semantics: one BDF integration per cell */
  PRX_REACTOR_ARKODE = 2     /* "ReactorArkode" PP/Source/Reactions/ReactorArkode.cpp:135-260 (ode.use_erkstep = 1, the default, or 0):
                                ARKODE's adaptive explicit Runge-Kutta stepper on the WHOLE BOX as one system - one step
                                size and one WRMS error norm for all cells of a react call, as the reference does on every
                                backend; FC = the box's RHS-evaluation count in every cell.  Array4 and 1-D react. */
} prx_reactor_kind;

/* cvode.solve_type (PP/Source/Reactions/ReactorCvode.cpp:507-934) */
typedef enum {
  PRX_SOLVE_DENSE_AJ_DIRECT = 0, /* "denseAJ_direct": analytic Jacobian + dense LU   */
  PRX_SOLVE_DENSE_DIRECT = 1     /* "dense_direct":   difference-quotient J + dense LU */
} prx_solve_type;

/* Filled by the shim from the same ParmParse keys the reference reads
 * (ode.* / cvode.*: ReactorCvode.cpp:18-35, ReactorRK64.cpp:12-18). */
typedef struct prx_config {
  int abi_version;          /* PRX_ABI_VERSION */
  int mechanism;            /* prx_mechanism */
  int reactor_kind;         /* prx_reactor_kind */
  int reactor_type;         /* 1 = e (UV, PeleC), 2 = h (HP, PeleLM): ReactorTypes.H:8-20 */
  double rtol;              /* ode.rtol  (CVODE only) */
  double atol;              /* ode.atol  */
  int max_order;            /* cvode.max_order (default 2, ReactorCvode.H:191) */
  int max_nls_iters;        /* ode.max_nls_iters (default 3) */
  int solve_type;           /* prx_solve_type */
  int rk64_nsubsteps_guess; /* ode.rk64_nsubsteps_guess (10)  */
  int rk64_nsubsteps_min;   /* ode.rk64_nsubsteps_min   (5)   */
  int rk64_nsubsteps_max;   /* ode.rk64_nsubsteps_max   (500) */
  int clean_init_massfrac;  /* ode.clean_init_massfrac (0) */
  int verbose;              /* ode.verbose */
  int device;               /* CUDA device ordinal, -1 = current device */
  int ark_method;           /* ode.rk_method (ReactorArkode.cpp:25-90): 40 Zonneveld 5-3-4 (default), 20 Heun-Euler, 30 Bogacki-
                               Shampine, 31 / 41 / 43 / 53 / 54 the ARK-L2SA explicit tables, 42 Sayfy-Aburub, 50 Cash-Karp,
                               51 Fehlberg 6-4-5, 52 Dormand-Prince, 60 Verner 8-5-6, 80 Fehlberg 13-7-8; any other value =
                               the default table, as in the reference */
  int ark_controller;       /* ode.rk_controller (ReactorArkode.cpp:90-116): 0 PID (default), 1 PI, 2 I */
  int ark_use_erkstep;      /* ode.use_erkstep (ReactorArkode.cpp:13): 1 = ERKStep (default), 0 = ARKStep with the same explicit
                               table - the same algorithm with ARKODE's generic adaptivity constants (safety 0.96, growth 20)
                               instead of ERKStep's (0.99, 25) */
} prx_config;

/* Same fields and meaning as amrex::Array4<T> (AMR/Src/Base/AMReX_Array4.H:60-68):
 * element (i,j,k,n) is p[(i-begin[0]) + (j-begin[1])*jstride + (k-begin[2])*kstride + n*nstride]. */
typedef struct prx_array4 {
  double* p;
  long long jstride, kstride, nstride;
  int begin[3];
  int end[3]; /* exclusive, as in AMReX */
  int ncomp;
} prx_array4;

typedef struct prx_array4i {
  int* p;
  long long jstride, kstride, nstride;
  int begin[3];
  int end[3];
  int ncomp;
} prx_array4i;

/* Work counters accumulated on the device since the last prx_reset_stats: the analogue of
 * CVODE's final statistics (ReactorCvode.cpp:1838-1901) and the input of the FLOP model
 * (SURVEY.md 8d).  Plain sums over cells so that ranks can be combined with one all-reduce. */
typedef struct prx_stats {
  long long cells;          /* cells integrated                                   */
  long long cells_masked;   /* cells skipped because mask == -1                   */
  long long n_steps;        /* internal steps (RK64 steps | CVODE nst)            */
  long long n_rhs;          /* RHS evaluations (RK64: 6/step | CVODE nfe + nfeDQ) */
  long long n_jac;          /* Jacobian evaluations (nje)                         */
  long long n_setups;       /* LU factorisations (nsetups)                        */
  long long n_newton;       /* Newton iterations = linear solves (nni)            */
  long long n_errfail;      /* error-test failures (netf)                         */
  long long n_convfail;     /* nonlinear convergence failures (ncfn)              */
  long long n_failed;       /* cells whose integrator returned a failure flag     */
  long long max_steps;      /* max internal steps over cells                      */
  long long reserved;
} prx_stats;

typedef struct prx_handle prx_handle;

/* ---- life cycle: ReactorBase::create + init / close (ReactorBase.H:20-23) ---- */
PRX_API int prx_create(const prx_config* cfg, prx_handle** out);
PRX_API int prx_destroy(prx_handle* h);
/* ReactorBase::set_typ_vals_ode (ReactorBase.cpp:6-31): n must be NUM_SPECIES+1 */
PRX_API int prx_set_typ_vals(prx_handle* h, const double* typ_vals_host, int n);
PRX_API const char* prx_last_error(const prx_handle* h);
PRX_API int prx_num_species(int mechanism);
PRX_API int prx_num_reactions(int mechanism);
PRX_API const char* prx_mechanism_name(int mechanism);
/* Mechanisms beyond the three above: any mechanism of PP/Mechanisms the generator pipeline reproduces can be compiled in
 * (pelec_b200/csrc/Makefile EXTRA_MECHS; ids 3, 4, ... in build order).  prx_mechanism_id: id by PelePhysics directory name,
 * PRX_ERR_UNSUPPORTED if this build does not hold it. */
PRX_API int prx_num_mechanisms(void);
PRX_API int prx_mechanism_id(const char* name);
PRX_API int prx_default_config(prx_config* cfg);

/* ---- ReactorBase::react, Array4 version (ReactorBase.H:25-40) ----
 * box = [lo, hi] inclusive.  rY in/out [N], rYsrc in [N], T in/out, rE in/out (+= dt*rEsrc),
 * rEsrc in, FC out (work count as double), mask in (-1 = skip; honoured by the CVODE reactor,
 * ignored by RK64, as in the reference).  Returns 0; the reference's return value (average
 * RK64 steps per cell / dummy 20) is available as prx_last_react_return(). */
PRX_API int prx_react_box(
  prx_handle* h, const int lo[3], const int hi[3], const prx_array4* rY, const prx_array4* rYsrc,
  const prx_array4* T, const prx_array4* rE, const prx_array4* rEsrc, const prx_array4* FC,
  const prx_array4i* mask, double dt, double time, void* cuda_stream);

/* ---- several boxes in one call: the caller's MFIter loop (PC/Source/React.cpp:133-138, ReactEval main.cpp:484-526) ----
 * lo / hi: [nboxes][3]; every array argument is an array of nboxes descriptors (mask may be NULL).  The boxes are independent;
 * they are issued on internal streams and run concurrently, and `cuda_stream` waits for all of them (stream-ordered like
 * prx_react_box: work queued on cuda_stream before the call completes first).  prx_last_react_return() covers all boxes. */
PRX_API int prx_react_boxes(
  prx_handle* h, int nboxes, const int* lo, const int* hi, const prx_array4* rY, const prx_array4* rYsrc,
  const prx_array4* T, const prx_array4* rE, const prx_array4* rEsrc, const prx_array4* FC, const prx_array4i* mask,
  double dt, double time, void* cuda_stream);

/* ---- PeleC::react_state, the per-box body (PC/Source/React.cpp:140-343), fused around the reactor ----
 * Replaces, for one box: the frcEExt kernel (React.cpp:181-220), the copies into STemp / extsrc_rY / extsrc_rE / fctCount /
 * dummyMask (React.cpp:95-119), reactor->react (:222-228), the unpack kernel (:233-321: S_new, I_R[0..N]) and the heat-release
 * kernel (:325-343: I_R[N+1]).  The integrator kernel reads S_old / S_new / non_react_src itself and writes S_new / I_R.
 *   sold   S_old (react_init: pass S_new)                 NVAR components, read only
 *   snew   S_new = S_old + dt*non_react_src on entry      NVAR components, updated iff do_update (= !react_init)
 *   nonrs  non-reacting source terms                      NVAR components, read only
 *   I_R    reaction source term, out                      NUM_SPECIES + 2 components (species, rho*E, heat release)
 *   work   optional (NULL): the reactor's FC per cell     1 component
 *   ix     component indices of the state vector (PC/Source/IndexDefines.H:3-12)
 * The reference's mask is 1 everywhere here (dummyMask, :178); EB-covered boxes are skipped by the caller (:151-158).  The
 * rotating-frame energy of pelec.do_rf: prx_set_rotating_frame below. */
typedef struct prx_state_index {
  int urho, umx, umy, umz, ueden, ueint, utemp, ufs, nvar;
} prx_state_index;

PRX_API int prx_react_state_box(
  prx_handle* h, const int lo[3], const int hi[3], const prx_array4* sold, const prx_array4* snew, const prx_array4* nonrs,
  const prx_array4* I_R, const prx_array4* work, const prx_state_index* ix, double dt, int do_update, void* cuda_stream);

/* Rotating frame of reference (pelec.do_rf, rf_omega, rf_axis, rf_axis_x/y/z; PC/Source/React.cpp:123-131,192-218,245-300):
 * the kinetic-energy bookkeeping of the react_state body then carries rotenrg = 0.5 omega^2 r^2 (PC/Source/Utilities.H:23-65)
 * with r the in-plane distance of the cell centre prob_lo + (index + 0.5) dx from the axis through axis_loc.  Box indices
 * passed to prx_react_state_box(es) are then the level's global cell indices (as AMReX boxes are).  on = 0 switches it off. */
PRX_API int prx_set_rotating_frame(
  prx_handle* h, int on, double omega, int axis, const double axis_loc[3], const double prob_lo[3], const double dx[3]);

PRX_API int prx_react_state_boxes(
  prx_handle* h, int nboxes, const int* lo, const int* hi, const prx_array4* sold, const prx_array4* snew,
  const prx_array4* nonrs, const prx_array4* I_R, const prx_array4* work, const prx_state_index* ix, double dt,
  int do_update, void* cuda_stream);

/* ---- PeleC::computeTemp's pointwise kernel (PC/Source/PeleC.cpp:2254-2272 -> pc_cmpTemp, PC/Source/Utilities.H:70-84) ----
 * state(UTEMP) <- REY2T(rho, rho*e_int / rho, rhoY / rho, guess = state(UTEMP)), in place, for every cell of every box; all
 * boxes on cuda_stream.  EOS only (no reactor state is used); reset_internal_energy stays with the caller.  SURVEY.md 8f-3. */
PRX_API int prx_compute_temp_boxes(
  prx_handle* h, int nboxes, const int* lo, const int* hi, const prx_array4* state, const prx_state_index* ix,
  void* cuda_stream);


/* ---- PeleC's conserved -> primitive conversion (PC/Source/Utilities.H:242-321 pc_ctoprim; called per box by
 * PeleC::getMOLSrcTerm / the Godunov driver, PC/Source/Hydro.cpp, PC/Source/MOL.cpp) ----
 * For every cell of every box: q = (rho, u, v, w, gamma_e, p, rho*e, T, Y_k), qaux = (gamma_c, c, c_small, dp/drho|e, dp/de|rho,
 * R_specific) from the conserved state u, with the Fuego EOS (T from (e, Y) by GET_T_GIVEN_EY with u(UTEMP) as the guess;
 * |Y_k| < 1e-4 eps flushed to zero as the reference does).  No rotating frame (omega = 0), NUM_ADV = NUM_AUX = NUM_LIN = 0.
 * DEVICE pointers, asynchronous on cuda_stream.  SURVEY.md 8f-3. */
typedef struct prx_prim_index {
  int qrho, qu, qv, qw, qgame, qpres, qreint, qtemp, qfs, qvar;   /* PC/Source/IndexDefines.H: QRHO ... QFS, QVAR */
  int qgamc, qc, qcsml, qdpdr, qdpde, qrspec, nqaux;              /* QGAMC ... QRSPEC, NQAUX */
} prx_prim_index;

PRX_API int prx_ctoprim_boxes(
  prx_handle* h, int nboxes, const int* lo, const int* hi, const prx_array4* u, const prx_array4* q, const prx_array4* qaux,
  const prx_state_index* ix, const prx_prim_index* qx, double small_num, void* cuda_stream);

/* ---- pele::physics::transport::SimpleTransport::get_transport_coeffs (PP/Source/Transport/Simple.H:391-447; PeleC's call
 * site PC/Source/Diffusion.cpp:182-222 with get_xi = get_mu = get_lam = get_Ddiag = true, get_chi = false) ----
 * Y [N], T [1], rho [1] in (component-offset views of the primitive FAB are fine); D [N] (the reference's rho*D_k, as stored
 * in coeff_cc(dComp_rhoD)), mu, xi, lam [1] out.  Fuego EOS (no Chung high-pressure corrections), Soret off. */
PRX_API int prx_transport_boxes(
  prx_handle* h, int nboxes, const int* lo, const int* hi, const prx_array4* Y, const prx_array4* T, const prx_array4* rho,
  const prx_array4* D, const prx_array4* mu, const prx_array4* xi, const prx_array4* lam, void* cuda_stream);


/* ---- ReactorBase::react, flat version (ReactorBase.H:42-54), DEVICE pointers ----
 * y [ncells*(N+1)] YCOrder (cell-major, T last), ysrc [ncells*N], e [ncells], esrc [ncells].
 * Reference semantics of the 1-D call: T is taken from y (no initial REY2T) and the
 * integrated T is returned (ReactorRK64.cpp:82-96,140-143). fc may be NULL.
 * ReactorCvode difference, by design: the reference's 1-D CVODE react (ReactorCvode.cpp:1491-1689) integrates the ncells
 * cells of ONE call as one coupled ODE system (one step size, one WRMS norm over all cells); this library integrates every
 * cell as its own system - the semantics of the reference's Array4 CPU path (ReactorCvode.cpp:1383-1464) - so results
 * coincide with the reference's for ncells = 1 and agree within the integrator tolerance otherwise.  The value the
 * reference returns from this call (nfe + nfeLS) is prx_last_react_return(): the largest per-cell count of the call.
 * ReactorArkode (ReactorArkode.cpp:266-380): the ncells cells of the call are ONE coupled ERK system here as in the
 * reference (one step size, one error norm); e advances by dt * esrc; prx_last_react_return() = nfe, also written to fc. */
PRX_API int prx_react_flat(
  prx_handle* h, double* y, const double* ysrc, double* e, const double* esrc, double* fc,
  double dt, double time, int ncells, void* cuda_stream);

/* ---- ReactorBase::flatten / unflatten (ReactorBase.H:56-80; FlattenOps, ReactorUtils.H:160-355) ----
 * DEVICE pointers.  ordering 0 = YCOrder (y[icell*(N+1)+k], the CPU/RK64 layout), 1 = CYOrder
 * (y[k*ncells+icell]) (ReactorUtils.H:50-76).  flatten also solves T from (e,Y) like box_flatten;
 * unflatten writes rhoY, rE = e + dt*rEsrc, T = REY2T(..), FC = fcunt (long int in the reference). */
PRX_API int prx_flatten(
  prx_handle* h, const int lo[3], const int hi[3], int ncells, int ordering, const prx_array4* rY,
  const prx_array4* rYsrc, const prx_array4* T, const prx_array4* rE, const prx_array4* rEsrc, double* y,
  double* ysrc, double* e, double* esrc, void* cuda_stream);
PRX_API int prx_unflatten(
  prx_handle* h, const int lo[3], const int hi[3], int ncells, int ordering, const prx_array4* rY,
  const prx_array4* T, const prx_array4* rE, const prx_array4* rEsrc, const prx_array4* FC, const double* y,
  const double* e, const long long* fcunt, double dt, void* cuda_stream);

/* ---- host-buffer convenience: one box held as SoA host arrays a[comp*ncells + cell]
 * (component stride = ncells, exactly an AMReX FAB); does H2D, react, D2H and synchronises.
 * This is the call the end-to-end benchmark and the ctypes tests use. */
PRX_API int prx_react_box_host(
  prx_handle* h, int nx, int ny, int nz, double* rY, const double* rYsrc, double* T, double* rE,
  const double* rEsrc, double* FC, const int* mask, double dt, double time);

/* flat react with HOST pointers: the reference's 1-D react takes host arrays even in GPU builds
 * (ReactorRK64.cpp:44-61,154-161 copies H2D/D2H itself). */
PRX_API int prx_react_flat_host(
  prx_handle* h, double* y, const double* ysrc, double* e, const double* esrc, double* fc, double dt,
  double time, int ncells);

/* ---- mechanism lookup for the C++ shim: the reference fixes the mechanism at compile time
 * (NUM_SPECIES / NUM_REACTIONS macros of mechanism.H, names from CKSYMS_STR); returns the
 * prx_mechanism id or PRX_ERR_UNSUPPORTED.  Name arguments may be NULL. */
PRX_API int prx_find_mechanism(int num_species, int num_reactions, const char* first_species, const char* last_species);
PRX_API const char* prx_species_name(int mechanism, int k);

/* ---- statistics ---- */
PRX_API int prx_reset_stats(prx_handle* h);
PRX_API int prx_get_stats(prx_handle* h, prx_stats* out); /* synchronises the handle's streams */
PRX_API int prx_last_react_return(prx_handle* h);        /* value ReactorBase::react would return */
PRX_API const char* prx_kernel_name(const prx_handle* h); /* name of the integrator kernel this handle launches */

/* ---- device probes used by the parity tests (n states, state-major host arrays) ---- */
PRX_API int prx_probe_rhs_host(
  prx_handle* h, int n, double t, const double* y, const double* rhoe_init, const double* rhoesrc,
  const double* rYsrc, double* ydot);
PRX_API int prx_probe_jac_host(prx_handle* h, int n, const double* y, double* J /* [n][(N+1)^2] col-major */);

/* ---- measurement helper: sustained FP64 FMA rate of the device [TFLOP/s] (register-resident DFMA
 * chains), the denominator of the FP64 roofline that bench.py reports ---- */
PRX_API int prx_measure_fp64_peak(int device, double* tflops_out);

#ifdef __cplusplus
}
#endif
#endif
