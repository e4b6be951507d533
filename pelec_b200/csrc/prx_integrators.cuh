// prx_integrators.cuh - per-cell time integrators written as RESUMABLE state machines.
//
// GPU design (why this is not a loop nest): one thread owns one cell, but the 32 cells of a
// warp are at different places of their adaptive integrations (different step counts, Newton
// iterations, Jacobian refreshes).  A plain transliteration of CVODE would serialise the warp on
// every expensive call site.  Here each integrator is a coroutine `resume(cell)` that runs the
// cheap control logic of ONE cell until it needs an expensive, uniform operation and then
// *yields a request*:
//     REQ_RHS  evaluate f(t, y)            (thermo + T-Newton + production rates)
//     REQ_JAC  evaluate the analytic Jacobian
//     REQ_LU   factor  A = I - gamma*J     (dense LU, partial pivoting)
//     REQ_DONE the cell is finished
// The kernel's warp loop (prx_kernels.cu) then serves all lanes' requests of one kind together
// at a single, converged code location, may defer rare kinds (JAC/LU) until enough lanes ask,
// and refills finished lanes from a global work queue (active-cell compaction).  Scheduling
// never changes per-cell arithmetic: each cell follows exactly the reference's sequence.
//
// Algorithms restated (SUN = PelePhysics/Submodules/sundials, PP = PelePhysics):
//   RK64  : ReactorRK64::react                     PP/Source/Reactions/ReactorRK64.cpp:166-320
//   BDF   : ReactorCvode::react CPU per-cell path  PP/Source/Reactions/ReactorCvode.cpp:1383-1464
//           CVode / cvHin / cvStep / cvNls ...     SUN/src/cvode/cvode.c:1030-1600, 2111-3665
//           Newton iteration + convergence test    SUN/src/sunnonlinsol/newton/sunnonlinsol_newton.c:176-330,
//                                                  SUN/src/cvode/cvode_nls.c:253-423
//           cvLsSetup / cvLsLinSys / cvLsSolve     SUN/src/cvode/cvode_ls.c:1308-1830
//           dense LU (GETRF/GETRS)                 SUN/src/sundials/sundials_dense.c:82-190
//           N_VLinearSum special cases             SUN/src/nvector/serial/nvector_serial.c
#ifndef PRX_INTEGRATORS_CUH
#define PRX_INTEGRATORS_CUH

#include <type_traits>

#include "prx_chem.cuh"

// Loop pragmas of the BDF control code.  The device build keeps these loops ROLLED: the control code is
// cheap arithmetic on shared-memory columns, and unrolling it (NEQ copies of every vector statement)
// is what pushed the first kernel's hot loop far beyond the instruction cache.
#ifndef PRX_VUNROLL_N
#define PRX_VUNROLL_N 1
#endif
#ifndef PRX_VLOOP
#ifdef PRX_HOST_TEST
#define PRX_VLOOP
#define PRX_MLOOP
#define PRX_ILOOP
#define PRX_UNROLL
#define PRX_ALOOP
#define PRX_QLOOP
#else
// one-thread-per-cell kernels with state in (L1/L2 backed) local memory: batches of independent loads
#define PRX_ALOOP _Pragma("unroll 10") // A = I - gamma*J: savedJ -> A, entries independent
#if PRX_VUNROLL_N > 1
#define PRX_QLOOP _Pragma("unroll")    // Nordsieck-array loops over the order, constant trip count + predicate
#else
#define PRX_QLOOP _Pragma("unroll 1")
#endif
#define PRX_UNROLL _Pragma("unroll")
#ifndef PRX_VUNROLL_N
#define PRX_VUNROLL_N 1
#endif
#define PRX_PRAGMA_(x) _Pragma(#x)
#define PRX_PRAGMA(x) PRX_PRAGMA_(x)
#define PRX_VLOOP PRX_PRAGMA(unroll PRX_VUNROLL_N) // per-component vector loops (-DPRX_VUNROLL_N=32: fully unrolled)
#define PRX_MLOOP _Pragma("unroll 1") // column loops of the dense matrix code
#define PRX_ILOOP _Pragma("unroll 2") // innermost row loops of the dense matrix code
#endif
#endif

// Per-component vector loop of the integrator: the storage policy decides which elements THIS thread owns.
//   serial policies (one thread per cell): all of them, i = 0, 1, ..., NEQ-1;
//   cooperative policy (G lanes per cell): i = lane_in_group, +G, ... - the loop body is one or two
//   SIMD statements and reductions go through s.gsum / s.gmax / s.gor (warp shuffles).
#define PRX_VEC_FOR(i) PRX_VLOOP for (int i = s.vbegin(); i < NEQ; i += s.vstep())

namespace prx {

// what a one-thread-per-cell storage policy answers to the cooperative hooks
struct SerialOwner
{
  PRX_MEM int vbegin() const { return 0; }
  PRX_MEM int vstep() const { return 1; }
  PRX_MEM bool owns(const int) const { return true; }
  PRX_MEM double gsum(const double x) const { return x; }
  PRX_MEM double gmax(const double x) const { return x; }
  PRX_MEM int gor(const int x) const { return x; }
  PRX_MEM void sync() const {}
};

enum Request { REQ_NONE = 0, REQ_RHS = 1, REQ_JAC = 2, REQ_LU = 3, REQ_DONE = 4 };

// which per-cell vector an RHS request reads / writes
enum VecSel { V_ZN0 = 0, V_ZN1 = 1, V_Y = 2, V_TEMPV = 3, V_FTEMP = 4 };

#define PRX_YIELD(LABEL, REQ) \
  do {                        \
    s.pc = LABEL;             \
    return REQ;               \
    case LABEL:;              \
  } while (0)

#define PRX_YIELD_RHS(LABEL, TVAL, INSEL, OUTSEL) \
  do {                                            \
    s.t_eval = (TVAL);                            \
    s.in_sel = (INSEL);                           \
    s.out_sel = (OUTSEL);                         \
    PRX_YIELD(LABEL, REQ_RHS);                    \
  } while (0)

// =====================================================================================
// RK64
// =====================================================================================
struct Rk64Params
{
  int reactor_type;
  double atol;
  int nsub_guess, nsub_min, nsub_max;
  int array4_semantics; // 1: re-solve T before/after (Array4 react); 0: flat react
};

struct Rk64Cell
{
  CellSrc src;
  double soln[NEQ], carry[NEQ], err[NEQ], ydot[NEQ];
  double current_time, time_init, time_out, dt_rk, dt_rk_min, dt_rk_max, dt_react;
  double t_eval;
  int nsteps, stage, pc;
  int in_sel, out_sel; // unused (RHS always soln -> ydot); kept for a uniform kernel interface
};

// ReactorRK64.H:9-40
PRX_DEVCONST double rk64_alpha[6] = {0.218150805229859, 0.256702469801519, 0.527402592007520,
                                     0.0484864267224467, 1.24517071533530, 0.412366034843237};
PRX_DEVCONST double rk64_beta[6] = {-0.113554138044166, -0.215118587818400, -0.0510152146250577,
                                    -1.07992686223881, -0.248664241213447, 0.0};
PRX_DEVCONST double rk64_err[6] = {-0.0554699315064507, 0.158481845574980, -0.0905918835751907,
                                   -0.219084567203338, 0.164022338959433, 0.0426421977505659};

// Start a cell: rY[N] in soln[0..N-1], T guess in soln[N], src filled by the caller.
PRX_DEV void rk64_begin(Rk64Cell& s, const Rk64Params& p, const double dt_react, const double time)
{
  s.time_init = time;
  s.time_out = time + dt_react;
  s.dt_react = dt_react;
  s.current_time = time;
  if (p.array4_semantics) {
    double rho = 0.0, Y[N];
    for (int k = 0; k < N; k++) {
      rho += s.soln[k];
    }
    const double rho_inv = 1.0 / rho;
    for (int k = 0; k < N; k++) {
      Y[k] = s.soln[k] * rho_inv;
    }
    double temp = s.soln[N];
    t_from_energy(p.reactor_type, s.src.rhoe_init * rho_inv, Y, temp);
    s.soln[N] = temp;
  }
  for (int k = 0; k < NEQ; k++) {
    s.carry[k] = s.soln[k];
  }
  s.dt_rk = dt_react / double(p.nsub_guess);
  s.dt_rk_min = dt_react / double(p.nsub_max);
  s.dt_rk_max = dt_react / double(p.nsub_min);
  s.nsteps = 0;
  s.pc = 0;
}

PRX_DEV int rk64_resume(Rk64Cell& s, const Rk64Params& p)
{
  switch (s.pc) {
  case 0:
    while (s.current_time < s.time_out) {
      for (int k = 0; k < NEQ; k++) {
        s.err[k] = 0.0;
      }
      for (s.stage = 0; s.stage < 6; s.stage++) {
        // NOTE (reference quirk 2): the step-start time is used for all six stages
        PRX_YIELD_RHS(1, s.current_time - s.time_init, V_Y, V_FTEMP);
        {
          const double ce = rk64_err[s.stage], ca = rk64_alpha[s.stage], cb = rk64_beta[s.stage];
          for (int k = 0; k < NEQ; k++) {
            s.err[k] += ce * s.dt_rk * s.ydot[k];
            s.soln[k] = s.carry[k] + ca * s.dt_rk * s.ydot[k];
            s.carry[k] = s.soln[k] + cb * s.dt_rk * s.ydot[k];
          }
        }
      }
      s.current_time += s.dt_rk;
      s.nsteps++;
      {
        double max_err = 1e-50;
        for (int k = 0; k < NEQ; k++) {
          const double a = fabs(s.err[k]);
          max_err = a > max_err ? a : max_err;
        }
        // a step is never rejected: the error estimate only steers the next dt (quirk 1)
        if (max_err < p.atol) {
          const double cf = 1.0 * pow(p.atol / max_err, 0.25);
          s.dt_rk = fmin(s.dt_rk_max, s.dt_rk * cf);
        } else {
          const double cf = 1.0 * pow(p.atol / max_err, 0.2);
          s.dt_rk = fmax(s.dt_rk_min, s.dt_rk * cf);
        }
        s.dt_rk = fmin(s.dt_rk, s.time_out - s.current_time);
      }
    }
    s.pc = 99;
  default:
    return REQ_DONE;
  }
}

// Finish a cell: returns rho*e_out; soln[N] becomes the output temperature.
PRX_DEV double rk64_finish(Rk64Cell& s, const Rk64Params& p)
{
  const double rhoe_out = s.src.rhoe_init + s.dt_react * s.src.rhoesrc;
  if (p.array4_semantics) {
    double rho = 0.0, Y[N];
    for (int k = 0; k < N; k++) {
      rho += s.soln[k];
    }
    const double rho_inv = 1.0 / rho;
    for (int k = 0; k < N; k++) {
      Y[k] = s.soln[k] * rho_inv;
    }
    double temp = s.soln[N];
    t_from_energy(p.reactor_type, rhoe_out * rho_inv, Y, temp);
    s.soln[N] = temp;
  }
  return rhoe_out;
}

// =====================================================================================
// Variable-order variable-step BDF (CVODE restated per cell)
// =====================================================================================
constexpr int QMAX_BDF = 5;
constexpr int LMAX = QMAX_BDF + 1;

struct BdfParams
{
  int reactor_type;
  double rtol;
  double atolv[NEQ]; // typ_vals[k]*atol or atol (ReactorUtils.H:119-141)
  int qmax;          // cvode.max_order
  int maxcor;        // ode.max_nls_iters
  int analytic_jac;  // denseAJ_direct (1) / dense_direct (0)
  int mxstep;        // 10000 (ReactorCvode.cpp:488)
  int maxnef;        // 100   (ReactorCvode.cpp:484)
  int msbj;          // 100   (ReactorCvode.cpp:497)
  int clean_init_massfrac;
  int array4_semantics;
};

enum BdfFlag {
  CV_SUCCESS = 0,
  CV_TOO_MUCH_WORK = -1,
  CV_TOO_MUCH_ACC = -2,
  CV_ERR_FAILURE = -3,
  CV_CONV_FAILURE = -4,
  CV_ILL_INPUT = -22,
  CV_TOO_CLOSE = -27
};

// ---- per-cell storage policies --------------------------------------------------------
// The integrator addresses its vectors through accessors so that the same state machine runs on
//  * BdfStoreLocal : plain arrays (host unit tests; legacy thread-local kernels), and
//  * BdfStoreSmem  : one shared-memory COLUMN per thread - element i of a vector lives at
//    base[(slot*NEQ + i) * STRIDE] with STRIDE = threads per block, so the 32 lanes of a warp touch
//    32 consecutive doubles (conflict-free) and the state never leaves the SM.  The zn slots are
//    sized by cvode.max_order at launch time; the linear-system matrix A optionally sits in the
//    same column (A_SMEM), savedJ (touched once per setup) stays in thread-local memory (L2).
template <int STRIDE>
struct SVecT
{
  double* p;
  PRX_MEM double& operator[](const int i) const { return p[i * STRIDE]; }
};

struct BdfStoreLocal : public SerialOwner
{
  typedef double* Vec;
  typedef double* Mat;
  double zn_[LMAX][NEQ], ewt_[NEQ], acor_[NEQ], tempv_[NEQ], ftemp_[NEQ], y_[NEQ], delta_[NEQ];
  double A_[NEQ * NEQ], savedJ_[NEQ * NEQ];
  int piv_[NEQ], perm_[NEQ];
  PRX_MEM int* perm() { return perm_; }
  PRX_MEM Vec zn(const int j) { return zn_[j]; }
  PRX_MEM Vec ewt() { return ewt_; }
  PRX_MEM Vec acor() { return acor_; }
  PRX_MEM Vec tempv() { return tempv_; }
  PRX_MEM Vec ftemp() { return ftemp_; }
  PRX_MEM Vec y() { return y_; }
  PRX_MEM Vec delta() { return delta_; }
  PRX_MEM Mat A() { return A_; }
  PRX_MEM double* savedJ() { return savedJ_; }
  PRX_MEM int* piv() { return piv_; }
  template <class MA, class VB>
  PRX_MEM void getrs(MA a, const int* p, VB b);
};

// The first generation's thread-local store with the vectors every Newton iteration reads (LEVEL >= 2: ewt, LEVEL >= 3: + acor)
// moved to shared-memory COLUMNS (element i at hot[i * STRIDE], STRIDE = threads per block): the block's otherwise unused
// shared memory holds what it can of the hot state at the first generation's occupancy; everything else stays in L1/L2-backed
// local memory.  LEVEL 1 is BdfStoreLocal (only the external source lives in shared memory, see CellSrcCol).
template <int STRIDE, int LEVEL>
struct BdfStoreLocalHot : public SerialOwner
{
  typedef double* Vec;
  typedef double* Mat;
  typedef SVecT<STRIDE> HVec;
  typedef typename std::conditional<(LEVEL >= 2), HVec, double*>::type EwtVec;
  typedef typename std::conditional<(LEVEL >= 3), HVec, double*>::type AcorVec;
  double* hot; // this thread's column: [ewt NEQ][acor NEQ]
  double zn_[LMAX][NEQ], tempv_[NEQ], ftemp_[NEQ], y_[NEQ], delta_[NEQ];
  double ewt_[(LEVEL >= 2) ? 1 : NEQ], acor_[(LEVEL >= 3) ? 1 : NEQ];
  double A_[NEQ * NEQ], savedJ_[NEQ * NEQ];
  int piv_[NEQ], perm_[NEQ];
  static PRX_HD int hot_doubles() { return ((LEVEL >= 2) ? NEQ : 0) + ((LEVEL >= 3) ? NEQ : 0); }
  PRX_MEM int* perm() { return perm_; }
  PRX_MEM Vec zn(const int j) { return zn_[j]; }
  PRX_MEM HVec ewt(std::true_type) { return HVec{hot}; }
  PRX_MEM double* ewt(std::false_type) { return ewt_; }
  PRX_MEM EwtVec ewt() { return ewt(std::integral_constant<bool, (LEVEL >= 2)>()); }
  PRX_MEM HVec acor(std::true_type) { return HVec{hot + NEQ * STRIDE}; }
  PRX_MEM double* acor(std::false_type) { return acor_; }
  PRX_MEM AcorVec acor() { return acor(std::integral_constant<bool, (LEVEL >= 3)>()); }
  PRX_MEM Vec tempv() { return tempv_; }
  PRX_MEM Vec ftemp() { return ftemp_; }
  PRX_MEM Vec y() { return y_; }
  PRX_MEM Vec delta() { return delta_; }
  PRX_MEM Mat A() { return A_; }
  PRX_MEM double* savedJ() { return savedJ_; }
  PRX_MEM int* piv() { return piv_; }
  template <class MA, class VB>
  PRX_MEM void getrs(MA a, const int* p, VB b);
};

template <int STRIDE, bool A_SMEM>
struct BdfStoreSmem : public SerialOwner
{
  typedef SVecT<STRIDE> Vec;
  typedef typename std::conditional<A_SMEM, SVecT<STRIDE>, double*>::type Mat;
  double* base; // this thread's column of the block's dynamic shared memory
  int nz;       // number of zn slots (qmax + 1)
  double A_[A_SMEM ? 1 : NEQ * NEQ], savedJ_[NEQ * NEQ];
  int piv_[NEQ], perm_[NEQ];
  PRX_MEM int* perm() { return perm_; }
  // vector slots after zn: ewt, acor, tempv, ftemp, y, delta, then N scratch doubles (chemistry), then A
  static PRX_HD int doubles_per_thread(const int qmax) { return (qmax + 1 + 6) * NEQ + N + (A_SMEM ? NEQ * NEQ : 0); }
  PRX_MEM Vec slot(const int s) const { return Vec{base + s * NEQ * STRIDE}; }
  PRX_MEM Vec zn(const int j) const { return slot(j); }
  PRX_MEM Vec ewt() const { return slot(nz); }
  PRX_MEM Vec acor() const { return slot(nz + 1); }
  PRX_MEM Vec tempv() const { return slot(nz + 2); }
  PRX_MEM Vec ftemp() const { return slot(nz + 3); }
  PRX_MEM Vec y() const { return slot(nz + 4); }
  PRX_MEM Vec delta() const { return slot(nz + 5); }
  PRX_MEM Vec scratch() const { return Vec{base + (nz + 6) * NEQ * STRIDE}; }
  PRX_MEM SVecT<STRIDE> A(std::true_type) const { return SVecT<STRIDE>{base + ((nz + 6) * NEQ + N) * STRIDE}; }
  PRX_MEM double* A(std::false_type) { return A_; }
  PRX_MEM Mat A() { return A(std::integral_constant<bool, A_SMEM>()); }
  PRX_MEM double* savedJ() { return savedJ_; }
  PRX_MEM int* piv() { return piv_; }
  template <class MA, class VB>
  PRX_MEM void getrs(MA a, const int* p, VB b);
};

// Hot/cold split, one cell per thread: the vectors every Newton iteration touches (zn0, zn1, ewt, acor, ftemp,
// y, delta [+ N species scratch]) are shared-memory columns, everything touched once per step or per linear
// setup (zn[2..], tempv, the Newton matrix A) is a column of a GLOBAL scratch buffer (coalesced: element i of a
// vector at cold[i*STRIDE], L2 resident), savedJ / pivots stay thread-local.  560 B of shared memory per LiDryer
// cell instead of 952 / 1752 -> 12 resident warps per SM instead of 7 / 4.
template <int STRIDE, bool SCRATCH>
struct BdfStoreHot : public SerialOwner
{
  typedef SVecT<STRIDE> Vec;
  typedef SVecT<STRIDE> Mat;
  double* hot;
  double* cold;
  int nz;
  double savedJ_[NEQ * NEQ];
  int piv_[NEQ], perm_[NEQ];
  PRX_MEM int* perm() { return perm_; }
  static PRX_HD int hot_doubles() { return 7 * NEQ + (SCRATCH ? N : 0); }
  static PRX_HD int cold_doubles(const int qmax) { return (qmax > 1 ? qmax - 1 : 0) * NEQ + NEQ + NEQ * NEQ; }
  PRX_MEM Vec hslot(const int s) const { return Vec{hot + s * NEQ * STRIDE}; }
  PRX_MEM Vec cslot(const int s) const { return Vec{cold + s * NEQ * STRIDE}; }
  PRX_MEM Vec zn(const int j) const { return (j < 2) ? hslot(j) : cslot(j - 2); }
  PRX_MEM Vec ewt() const { return hslot(2); }
  PRX_MEM Vec acor() const { return hslot(3); }
  PRX_MEM Vec ftemp() const { return hslot(4); }
  PRX_MEM Vec y() const { return hslot(5); }
  PRX_MEM Vec delta() const { return hslot(6); }
  PRX_MEM Vec scratch() const { return hslot(7); }
  PRX_MEM Vec tempv() const { return cslot(nz - 2); }
  PRX_MEM Mat A() const { return cslot(nz - 1); }
  PRX_MEM double* savedJ() { return savedJ_; }
  PRX_MEM int* piv() { return piv_; }
  template <class MA, class VB>
  PRX_MEM void getrs(MA a, const int* p, VB b);
};

template <class Store, class Src = CellSrc>
struct BdfCellT : public Store
{
  typedef typename Store::Vec Vec;
  Src src;
  // step / order control
  double tn, tout, t0, h, hprime, hscale, eta, etamax, hu;
  double gamma, gammap, gamrat, rl1, crate, delp, acnrm, dsm, saved_t;
  double tq[6], l[LMAX], tau[LMAX + 1];
  double saved_tq5;
  double hg, hub, hlb, hnew; // cvHin
  double minInc, yjsaved, inc; // difference-quotient Jacobian
  double t_eval;
  int q, qprime, L, qwait, qu, indx_acor, jcur;
  int nflag, ncf, nef, convfail, callSetup, jbad_nls, nls_jcur, curiter, retval, nls_ok;
  int count1, jcol, lu_info, flag;
  int nst, nfe, nfeDQ, ncfn, netf, nni, nsetups, nstlp, nje, nstlj, nstloc;
  int pc, in_sel, out_sel;
  PRX_MEM Vec vec(const int sel)
  {
    switch (sel) {
    case V_ZN0:
      return this->zn(0);
    case V_ZN1:
      return this->zn(1);
    case V_Y:
      return this->y();
    case V_TEMPV:
      return this->tempv();
    default:
      return this->ftemp();
    }
  }
};
typedef BdfCellT<BdfStoreLocal> BdfCell;

constexpr double UROUND = 2.220446049250313e-16;
constexpr int FIRST_CALL = 6, PREV_CONV_FAIL = 7, PREV_ERR_FAIL = 8;
constexpr int NLS_CONTINUE = 901, NLS_CONV_RECVR = 902;

template <class Cell, class VX, class VW>
PRX_DEV double wrms(Cell& s, VX x, VW w)
{
  double sum = 0.0;
  PRX_VEC_FOR(i) {
    const double pr = x[i] * w[i];
    sum += pr * pr;
  }
  sum = s.gsum(sum);
  return sqrt(sum / NEQ);
}

template <class Cell, class VY>
PRX_DEV int ewt_set(Cell& s, const BdfParams& p, VY ycur)
{
  int bad = 0;
  PRX_VEC_FOR(i) {
    const double v = p.rtol * fabs(ycur[i]) + p.atolv[i];
    bad |= !(v > 0.0);
    s.tempv()[i] = v;
  }
  bad = s.gor(bad);
  if (bad) {
    return -1;
  }
  PRX_VEC_FOR(i) {
    s.ewt()[i] = 1.0 / s.tempv()[i];
  }
  return 0;
}

PRX_DEV double rpowerR(const double base, const double e) { return (base <= 0.0) ? 0.0 : pow(base, e); }
PRX_DEV double rpowerI(const double base, const int e)
{
  double prod = 1.0;
  const int expt = e < 0 ? -e : e;
  for (int i = 1; i <= expt; i++) {
    prod *= base;
  }
  return (e < 0) ? 1.0 / prod : prod;
}

template <class Cell>
PRX_DEV void bdf_rescale(Cell& s)
{
  double factor = s.eta;
  for (int j = 1; j <= s.q; j++) {
    PRX_VEC_FOR(i) {
      s.zn(j)[i] *= factor;
    }
    factor *= s.eta;
  }
  s.h = s.hscale * s.eta;
  s.hscale = s.h;
}

template <class Cell>
PRX_DEV void bdf_increase_order(Cell& s, const int qmax)
{
  double alpha0, alpha1, prod, xi, xiold, hsum;
  for (int i = 0; i <= qmax; i++) {
    s.l[i] = 0.0;
  }
  s.l[2] = alpha1 = prod = xiold = 1.0;
  alpha0 = -1.0;
  hsum = s.hscale;
  if (s.q > 1) {
    for (int j = 1; j < s.q; j++) {
      hsum += s.tau[j + 1];
      xi = hsum / s.hscale;
      prod *= xi;
      alpha0 -= 1.0 / (j + 1);
      alpha1 += 1.0 / xi;
      for (int i = j + 2; i >= 2; i--) {
        s.l[i] = s.l[i] * xiold + s.l[i - 1];
      }
      xiold = xi;
    }
  }
  const double A1 = (-alpha0 - alpha1) / prod;
  PRX_VEC_FOR(i) {
    s.zn(s.L)[i] = A1 * s.zn(s.indx_acor)[i];
  }
  if (s.q > 1) {
    for (int j = 2; j <= s.q; j++) {
      PRX_VEC_FOR(i) {
        s.zn(j)[i] += s.l[j] * s.zn(s.L)[i];
      }
    }
  }
}

template <class Cell>
PRX_DEV void bdf_decrease_order(Cell& s, const int qmax)
{
  for (int i = 0; i <= qmax; i++) {
    s.l[i] = 0.0;
  }
  s.l[2] = 1.0;
  double hsum = 0.0;
  for (int j = 1; j <= s.q - 2; j++) {
    hsum += s.tau[j];
    const double xi = hsum / s.hscale;
    for (int i = j + 2; i >= 2; i--) {
      s.l[i] = s.l[i] * xi + s.l[i - 1];
    }
  }
  if (s.q > 2) {
    for (int j = 2; j < s.q; j++) {
      PRX_VEC_FOR(i) {
        s.zn(j)[i] += -s.l[j] * s.zn(s.q)[i];
      }
    }
  }
}

template <class Cell>
PRX_DEV void bdf_adjust_order(Cell& s, const int deltaq, const int qmax)
{
  if ((s.q == 2) && (deltaq != 1)) {
    return;
  }
  if (deltaq == 1) {
    bdf_increase_order(s, qmax);
  } else if (deltaq == -1) {
    bdf_decrease_order(s, qmax);
  }
}

// cvPredict / cvRestore (cvode.c:2687-2717, 3386-3401): zn[j-1] +/-= zn[j] for k = 1..q, j = q..k.  Per component the q+1
// Nordsieck entries are loaded once, the triangle of additions runs in registers in the reference's order, and q entries are
// stored: 2q+1 memory accesses per component instead of 3q(q+1)/2 (ncu r02a: the statement-by-statement form was 11.8 % of
// the gen-1 kernel's local-memory traffic).  Measured on B200 (session r02f): SLOWER, 70.6 vs 67.0 ms per 128^3 LiDryer - the
// statement form keeps 10 independent loads in flight per statement, the register form serialises them behind its triangle;
// kept behind PRX_PASCAL_PREDICT (off) and bit-identical (tests/test_host_logic.py passes with either).
template <class Cell, int SIGN>
PRX_DEV void bdf_pascal(Cell& s)
{
  const int q = s.q;
  PRX_VEC_FOR(i) {
    double z[LMAX];
    PRX_UNROLL
    for (int j = 0; j < LMAX; j++) {
      z[j] = 0.0;
      if (j <= q) {
        z[j] = s.zn(j)[i];
      }
    }
    PRX_UNROLL
    for (int k = 1; k < LMAX; k++) {
      PRX_UNROLL
      for (int j = LMAX - 1; j >= k; j--) {
        if (j <= q) { // (k <= j)
          z[j - 1] = (SIGN > 0) ? (z[j - 1] + z[j]) : (z[j - 1] - z[j]);
        }
      }
    }
    PRX_UNROLL
    for (int j = 0; j < LMAX - 1; j++) {
      if (j < q) {
        s.zn(j)[i] = z[j];
      }
    }
  }
}

#ifndef PRX_PASCAL_PREDICT
#define PRX_PASCAL_PREDICT 0 // measured (session r02f, 128^3 LiDryer): 70.6 ms with the register form vs 67.0 ms with the statement form
#endif
template <class Cell>
PRX_DEV void bdf_predict(Cell& s)
{
  s.tn += s.h;
#if PRX_PASCAL_PREDICT
  bdf_pascal<Cell, 1>(s);
#else
  for (int k = 1; k <= s.q; k++) {
    for (int j = s.q; j >= k; j--) {
      PRX_VEC_FOR(i) {
        s.zn(j - 1)[i] = s.zn(j - 1)[i] + s.zn(j)[i];
      }
    }
  }
#endif
}

template <class Cell>
PRX_DEV void bdf_restore(Cell& s)
{
  s.tn = s.saved_t;
#if PRX_PASCAL_PREDICT
  bdf_pascal<Cell, -1>(s);
#else
  for (int k = 1; k <= s.q; k++) {
    for (int j = s.q; j >= k; j--) {
      PRX_VEC_FOR(i) {
        s.zn(j - 1)[i] = s.zn(j - 1)[i] - s.zn(j)[i];
      }
    }
  }
#endif
}

// cvSet -> cvSetBDF -> cvSetTqBDF (cvode.c:2730-2999)
template <class Cell>
PRX_DEV void bdf_set(Cell& s)
{
  double alpha0, alpha0_hat, xi_inv, xistar_inv, hsum;
  const int q = s.q;
  s.l[0] = s.l[1] = xi_inv = xistar_inv = 1.0;
  for (int i = 2; i <= q; i++) {
    s.l[i] = 0.0;
  }
  alpha0 = alpha0_hat = -1.0;
  hsum = s.h;
  if (q > 1) {
    for (int j = 2; j < q; j++) {
      hsum += s.tau[j - 1];
      xi_inv = s.h / hsum;
      alpha0 -= 1.0 / j;
      for (int i = j; i >= 1; i--) {
        s.l[i] += s.l[i - 1] * xi_inv;
      }
    }
    alpha0 -= 1.0 / q;
    xistar_inv = -s.l[1] - alpha0;
    hsum += s.tau[q - 1];
    xi_inv = s.h / hsum;
    alpha0_hat = -s.l[1] - xi_inv;
    for (int i = q; i >= 1; i--) {
      s.l[i] += s.l[i - 1] * xistar_inv;
    }
  }
  {
    const double A1 = 1.0 - alpha0_hat + alpha0;
    const double A2 = 1.0 + q * A1;
    s.tq[2] = fabs(A1 / (alpha0 * A2));
    s.tq[5] = fabs(A2 * xistar_inv / (s.l[q] * xi_inv));
    if (s.qwait == 1) {
      if (q > 1) {
        const double C = xistar_inv / s.l[q];
        const double A3 = alpha0 + 1.0 / q;
        const double A4 = alpha0_hat + xi_inv;
        const double Cpinv = (1.0 - A4 + A3) / A3;
        s.tq[1] = fabs(C * Cpinv);
      } else {
        s.tq[1] = 1.0;
      }
      hsum += s.tau[q];
      xi_inv = s.h / hsum;
      const double A5 = alpha0 - (1.0 / (q + 1));
      const double A6 = alpha0_hat - xi_inv;
      const double Cppinv = (1.0 - A6 + A5) / A2;
      s.tq[3] = fabs(Cppinv / (xi_inv * (q + 2) * A5));
    }
    s.tq[4] = 0.1 / s.tq[2];
  }
  s.rl1 = 1.0 / s.l[1];
  s.gamma = s.h * s.rl1;
  if (s.nst == 0) {
    s.gammap = s.gamma;
  }
  s.gamrat = (s.nst > 0) ? s.gamma / s.gammap : 1.0;
}

// Dense LU with partial pivoting, same pivot choice and operation order as
// SUNDlsMat_denseGETRF (column-major, row interchanges applied to the whole row).
template <class MA>
PRX_DEV int dense_getrf(MA a, int* p)
{
  PRX_MLOOP
  for (int k = 0; k < NEQ; k++) {
    const int ck = k * NEQ;
    int l = k;
    PRX_VLOOP
    for (int i = k + 1; i < NEQ; i++) {
      if (fabs(a[ck + i]) > fabs(a[ck + l])) {
        l = i;
      }
    }
    p[k] = l;
    if (a[ck + l] == 0.0) {
      return k + 1;
    }
    if (l != k) {
      PRX_VLOOP
      for (int i = 0; i < NEQ; i++) {
        const double temp = a[i * NEQ + l];
        a[i * NEQ + l] = a[i * NEQ + k];
        a[i * NEQ + k] = temp;
      }
    }
    const double mult = 1.0 / a[ck + k];
    PRX_ILOOP
    for (int i = k + 1; i < NEQ; i++) {
      a[ck + i] *= mult;
    }
    PRX_MLOOP
    for (int j = k + 1; j < NEQ; j++) {
      const int cj = j * NEQ;
      const double a_kj = a[cj + k];
      if (a_kj != 0.0) {
        PRX_ILOOP
        for (int i = k + 1; i < NEQ; i++) {
          a[cj + i] -= a_kj * a[ck + i];
        }
      }
    }
  }
  return 0;
}

template <class MA, class VB>
PRX_DEV void dense_getrs(MA a, const int* p, VB b)
{
  PRX_VLOOP
  for (int k = 0; k < NEQ; k++) {
    const int pk = p[k];
    if (pk != k) {
      const double tmp = b[k];
      b[k] = b[pk];
      b[pk] = tmp;
    }
  }
  PRX_MLOOP
  for (int k = 0; k < NEQ - 1; k++) {
    const int ck = k * NEQ;
    const double bk = b[k];
    PRX_ILOOP
    for (int i = k + 1; i < NEQ; i++) {
      b[i] -= a[ck + i] * bk;
    }
  }
  PRX_MLOOP
  for (int k = NEQ - 1; k > 0; k--) {
    const int ck = k * NEQ;
    b[k] /= a[ck + k];
    const double bk = b[k];
    PRX_ILOOP
    for (int i = 0; i < k; i++) {
      b[i] -= a[ck + i] * bk;
    }
  }
  b[0] /= a[0];
}

// Device forms of GETRF / GETRS for small systems, one thread per matrix, the matrix in (L2-backed) memory.
// ncu (profiles/r01d_*): with rolled loops every a[...] access of the elimination was a dependent memory round
// trip (long-scoreboard stalls = 45-49 % of all samples, concentrated here).  These forms do the same
// per-entry arithmetic in the same order, but with compile-time row/column offsets and all loads of a step
// issued before its stores, so the accesses of one elimination step are in flight together.
template <class MA>
PRX_DEV int dense_getrf_small(MA a, int* p, int* perm = nullptr)
{
  if (perm != nullptr) { // composed row permutation, so that a solve gathers b in one step (see dense_getrs_small)
    PRX_UNROLL
    for (int i = 0; i < NEQ; i++) {
      perm[i] = i;
    }
  }
  PRX_MLOOP
  for (int k = 0; k < NEQ; k++) {
    const int ck = k * NEQ;
    double col[NEQ];
PRX_UNROLL
    for (int i = 0; i < NEQ; i++) {
      col[i] = (i >= k) ? a[ck + i] : 0.0;
    }
    int l = k;
    double best = -1.0, pv = 0.0, akk = 0.0;
PRX_UNROLL
    for (int i = 0; i < NEQ; i++) {
      if (i >= k) {
        const double v = fabs(col[i]);
        if (i == k) {
          akk = col[i];
          pv = col[i];
          best = v;
        } else if (v > best) {
          best = v;
          l = i;
          pv = col[i];
        }
      }
    }
    p[k] = l;
    if (perm != nullptr && l != k) {
      const int t = perm[k];
      perm[k] = perm[l];
      perm[l] = t;
    }
    if (pv == 0.0) {
      return k + 1;
    }
    double rk[NEQ]; // row k of every column after the interchange
PRX_UNROLL
    for (int j = 0; j < NEQ; j++) {
      rk[j] = a[j * NEQ + l];
    }
    if (l != k) {
      double t[NEQ];
PRX_UNROLL
      for (int j = 0; j < NEQ; j++) {
        t[j] = a[j * NEQ + k];
      }
PRX_UNROLL
      for (int j = 0; j < NEQ; j++) {
        a[j * NEQ + l] = t[j];
        a[j * NEQ + k] = rk[j];
      }
    }
    const double mult = 1.0 / pv;
    double m[NEQ];
PRX_UNROLL
    for (int i = 0; i < NEQ; i++) {
      const double ci = (i == l) ? akk : col[i]; // column k below the diagonal, after the interchange
      m[i] = ci * mult;
      if (i > k) {
        a[ck + i] = m[i];
      }
    }
PRX_UNROLL
    for (int j = 0; j < NEQ; j++) {
      if (j > k) {
        const double a_kj = rk[j];
        double x[NEQ];
PRX_UNROLL
        for (int i = 0; i < NEQ; i++) {
          x[i] = (i > k) ? a[j * NEQ + i] : 0.0;
        }
        if (a_kj != 0.0) {
PRX_UNROLL
          for (int i = 0; i < NEQ; i++) {
            if (i > k) {
              a[j * NEQ + i] = x[i] - a_kj * m[i];
            }
          }
        }
      }
    }
  }
  return 0;
}

// Same, for mid-size systems (NEQ <= 24): the column loop stays rolled (one memory round trip per column
// instead of one per entry), rows are compile-time offsets; row interchanges in chunks of 8 columns.
template <class MA>
PRX_DEV int dense_getrf_mid(MA a, int* p, int* perm = nullptr)
{
  if (perm != nullptr) {
    PRX_UNROLL
    for (int i = 0; i < NEQ; i++) {
      perm[i] = i;
    }
  }
  PRX_MLOOP
  for (int k = 0; k < NEQ; k++) {
    const int ck = k * NEQ;
    double m[NEQ];
    int l = k;
    double best = -1.0, pv = 0.0, akk = 0.0;
    PRX_UNROLL
    for (int i = 0; i < NEQ; i++) {
      m[i] = (i >= k) ? a[ck + i] : 0.0;
    }
    PRX_UNROLL
    for (int i = 0; i < NEQ; i++) {
      if (i >= k) {
        const double v = fabs(m[i]);
        if (i == k) {
          akk = m[i];
          pv = m[i];
          best = v;
        } else if (v > best) {
          best = v;
          l = i;
          pv = m[i];
        }
      }
    }
    p[k] = l;
    if (perm != nullptr && l != k) {
      const int t = perm[k];
      perm[k] = perm[l];
      perm[l] = t;
    }
    if (pv == 0.0) {
      return k + 1;
    }
    if (l != k) {
      PRX_MLOOP
      for (int j0 = 0; j0 < NEQ; j0 += 8) {
        double tl[8], tk[8];
        PRX_UNROLL
        for (int jj = 0; jj < 8; jj++) {
          const int j = (j0 + jj < NEQ) ? j0 + jj : NEQ - 1;
          tl[jj] = a[j * NEQ + l];
          tk[jj] = a[j * NEQ + k];
        }
        PRX_UNROLL
        for (int jj = 0; jj < 8; jj++) {
          if (j0 + jj < NEQ) {
            a[(j0 + jj) * NEQ + l] = tk[jj];
            a[(j0 + jj) * NEQ + k] = tl[jj];
          }
        }
      }
    }
    const double mult = 1.0 / pv;
    PRX_UNROLL
    for (int i = 0; i < NEQ; i++) {
      const double ci = (i == l) ? akk : m[i];
      m[i] = ci * mult;
      if (i > k) {
        a[ck + i] = m[i];
      }
    }
    PRX_MLOOP
    for (int j = k + 1; j < NEQ; j++) {
      const int cj = j * NEQ;
      const double a_kj = a[cj + k];
      double x[NEQ];
      PRX_UNROLL
      for (int i = 0; i < NEQ; i++) {
        x[i] = (i > k) ? a[cj + i] : 0.0;
      }
      if (a_kj != 0.0) {
        PRX_UNROLL
        for (int i = 0; i < NEQ; i++) {
          if (i > k) {
            a[cj + i] = x[i] - a_kj * m[i];
          }
        }
      }
    }
  }
  return 0;
}

template <class MA, class VB>
PRX_DEV void dense_getrs_small(MA a, const int* p, VB b, const int* perm = nullptr)
{
  double bb[NEQ];
  if (perm != nullptr) { // the NEQ row interchanges at once: independent gathers instead of a dependent swap chain
    PRX_UNROLL
    for (int i = 0; i < NEQ; i++) {
      bb[i] = b[perm[i]];
    }
  } else {
    PRX_MLOOP
    for (int k = 0; k < NEQ; k++) {
      const int pk = p[k];
      if (pk != k) {
        const double tmp = b[k];
        b[k] = b[pk];
        b[pk] = tmp;
      }
    }
    PRX_UNROLL
    for (int i = 0; i < NEQ; i++) {
      bb[i] = b[i];
    }
  }
PRX_UNROLL
  for (int k = 0; k < NEQ - 1; k++) {
PRX_UNROLL
    for (int i = k + 1; i < NEQ; i++) {
      bb[i] -= a[k * NEQ + i] * bb[k];
    }
  }
PRX_UNROLL
  for (int k = NEQ - 1; k > 0; k--) {
    bb[k] /= a[k * NEQ + k];
PRX_UNROLL
    for (int i = 0; i < k; i++) {
      bb[i] -= a[k * NEQ + i] * bb[k];
    }
  }
  bb[0] /= a[0];
PRX_UNROLL
  for (int i = 0; i < NEQ; i++) {
    b[i] = bb[i];
  }
}

template <class MA, class VB>
PRX_MEM void BdfStoreLocal::getrs(MA a, const int* p, VB b)
{
#ifndef PRX_HOST_TEST
  if (NEQ <= 24) {
    dense_getrs_small(a, p, b, perm_);
    return;
  }
#endif
  dense_getrs(a, p, b);
}
template <int STRIDE, int LEVEL>
template <class MA, class VB>
PRX_MEM void BdfStoreLocalHot<STRIDE, LEVEL>::getrs(MA a, const int* p, VB b)
{
  if (NEQ <= 24) {
    dense_getrs_small(a, p, b, perm_);
    return;
  }
  dense_getrs(a, p, b);
}

template <int STRIDE, bool A_SMEM>
template <class MA, class VB>
PRX_MEM void BdfStoreSmem<STRIDE, A_SMEM>::getrs(MA a, const int* p, VB b)
{
#ifndef PRX_HOST_TEST
  if (NEQ <= 24) {
    dense_getrs_small(a, p, b, perm_);
    return;
  }
#endif
  dense_getrs(a, p, b);
}

template <int STRIDE, bool SCRATCH>
template <class MA, class VB>
PRX_MEM void BdfStoreHot<STRIDE, SCRATCH>::getrs(MA a, const int* p, VB b)
{
#ifndef PRX_HOST_TEST
  if (NEQ <= 24) {
    dense_getrs_small(a, p, b, perm_);
    return;
  }
#endif
  dense_getrs(a, p, b);
}

template <class Cell>
PRX_DEV void bdf_set_eta(Cell& s)
{
  if ((s.eta > 0.0) && (s.eta < 1.5)) {
    s.eta = 1.0;
    s.hprime = s.h;
  } else {
    if (s.eta >= 1.5) {
      s.eta = fmin(s.eta, s.etamax);
      s.eta /= fmax(1.0, fabs(s.h) * 0.0 * s.eta);
    } else {
      s.eta = fmax(s.eta, 0.1);
      s.eta = fmax(s.eta, 0.0 / fabs(s.h));
    }
    s.hprime = s.h * s.eta;
  }
}

// cvPrepareNextStep (+cvComputeEtaqm1/qp1, cvChooseEta), cvode.c:3468-3665
template <class Cell>
PRX_DEV void bdf_prepare_next_step(Cell& s, const int qmax)
{
  if (s.etamax == 1.0) {
    s.qwait = (s.qwait > 2) ? s.qwait : 2;
    s.qprime = s.q;
    s.hprime = s.h;
    s.eta = 1.0;
    return;
  }
  const double etaq = 1.0 / (rpowerR(6.0 * s.dsm, 1.0 / s.L) + 0.000001);
  if (s.qwait != 0) {
    s.eta = etaq;
    s.qprime = s.q;
    bdf_set_eta(s);
    return;
  }
  s.qwait = 2;
  double etaqm1 = 0.0;
  if (s.q > 1) {
    const double ddn = wrms(s, s.zn(s.q), s.ewt()) * s.tq[1];
    etaqm1 = 1.0 / (rpowerR(6.0 * ddn, 1.0 / s.q) + 0.000001);
  }
  double etaqp1 = 0.0;
  if (s.q != qmax) {
    if (s.saved_tq5 != 0.0) {
      const double cquot = (s.tq[5] / s.saved_tq5) * rpowerI(s.h / s.tau[2], s.L);
      PRX_VEC_FOR(i) {
        s.tempv()[i] = -cquot * s.zn(qmax)[i] + s.acor()[i];
      }
      const double dup = wrms(s, s.tempv(), s.ewt()) * s.tq[3];
      etaqp1 = 1.0 / (rpowerR(10.0 * dup, 1.0 / (s.L + 1)) + 0.000001);
    }
  }
  const double etam = fmax(etaqm1, fmax(etaq, etaqp1));
  if ((etam > 0.0) && (etam < 1.5)) {
    s.eta = 1.0;
    s.qprime = s.q;
  } else if (etam == etaq) {
    s.eta = etaq;
    s.qprime = s.q;
  } else if (etam == etaqm1) {
    s.eta = etaqm1;
    s.qprime = s.q - 1;
  } else {
    s.eta = etaqp1;
    s.qprime = s.q + 1;
    PRX_VEC_FOR(i) {
      s.zn(qmax)[i] = s.acor()[i];
    }
  }
  bdf_set_eta(s);
}

// Start a cell: zn[0] = (rhoY, T guess); src filled by the caller. (box_flatten, ReactorUtils.H:160-220)
template <class Cell>
PRX_DEV void bdf_begin(Cell& s, const BdfParams& p, const double dt_react, const double time)
{
  if (p.array4_semantics) {
    double rho = 0.0, Y[N];
    for (int k = 0; k < N; k++) {
      rho += s.zn(0)[k];
    }
    const double rho_inv = 1.0 / rho;
    for (int k = 0; k < N; k++) {
      Y[k] = s.zn(0)[k] * rho_inv;
    }
    double temp = s.zn(0)[N];
    t_from_energy(p.reactor_type, s.src.rhoe_init * rho_inv, Y, temp);
    s.zn(0)[N] = temp;
  }
  s.t0 = time;
  s.tout = time + dt_react;
  s.pc = 0;
}

// The CVODE state machine.  Returns the next request; on REQ_DONE the answer is in s.y().
template <class Cell>
PRX_DEV int bdf_resume(Cell& s, const BdfParams& p)
{
  switch (s.pc) {
  case 0:
    // ---- CVodeReInit (cvode.c) ----
    s.tn = s.t0;
    s.q = 1;
    s.L = 2;
    s.qwait = s.L;
    s.etamax = 10000.0;
    s.qu = 0;
    s.hu = 0.0;
    s.nst = s.nfe = s.nfeDQ = s.ncfn = s.netf = s.nni = s.nsetups = s.nstlp = 0;
    s.nje = 0;
    s.nstlj = 0;
    s.flag = CV_SUCCESS;
    s.saved_tq5 = 0.0;
    s.indx_acor = p.qmax;
    s.crate = 1.0;
    s.delp = 0.0;
    s.gammap = 0.0;
    for (int i = 0; i < LMAX + 1; i++) {
      s.tau[i] = 0.0;
    }
    // ---- first call block of CVode ----
    if (ewt_set(s, p, s.zn(0)) != 0) {
      s.flag = CV_ILL_INPUT;
      PRX_VEC_FOR(i) {
        s.y()[i] = s.zn(0)[i];
      }
      break;
    }
    PRX_YIELD_RHS(1, s.tn, V_ZN0, V_ZN1);
    s.nfe++;
    // ---- cvHin ----
    {
      const double tdist = fabs(s.tout - s.tn);
      const double tround = UROUND * fmax(fabs(s.tn), fabs(s.tout));
      if (s.tout == s.tn || tdist < 2.0 * tround) {
        s.flag = CV_TOO_CLOSE;
        PRX_VEC_FOR(i) {
          s.y()[i] = s.zn(0)[i];
        }
        break;
      }
      s.hlb = 100.0 * tround;
      double hub_inv = 0.0;
      PRX_VEC_FOR(i) {
        const double t2 = fabs(s.zn(0)[i]);
        const double t1 = 0.1 * t2 + 1.0 / s.ewt()[i];
        const double r = fabs(s.zn(1)[i]) / t1;
        hub_inv = (r > hub_inv) ? r : hub_inv;
      }
      hub_inv = s.gmax(hub_inv);
      s.hub = 0.1 * tdist;
      if (s.hub * hub_inv > 1.0) {
        s.hub = 1.0 / hub_inv;
      }
      s.hg = sqrt(s.hlb * s.hub);
    }
    if (s.hub < s.hlb) {
      s.h = s.hg;
    } else {
      s.hnew = s.hg;
      for (s.count1 = 1; s.count1 <= 4; s.count1++) {
        PRX_VEC_FOR(i) {
          s.y()[i] = s.hg * s.zn(1)[i] + s.zn(0)[i];
        }
        PRX_YIELD_RHS(2, s.tn + s.hg, V_Y, V_TEMPV);
        s.nfe++;
        {
          const double ihg = 1.0 / s.hg;
          PRX_VEC_FOR(i) {
            s.tempv()[i] = ihg * (s.tempv()[i] - s.zn(1)[i]); // N_VLinearSum a == -b: a*(x-y)
          }
          const double yddnrm = wrms(s, s.tempv(), s.ewt());
          s.hnew = (yddnrm * s.hub * s.hub > 2.0) ? sqrt(2.0 / yddnrm) : sqrt(s.hg * s.hub);
          if (s.count1 == 4) {
            break;
          }
          const double hrat = s.hnew / s.hg;
          if ((hrat > 0.5) && (hrat < 2.0)) {
            break;
          }
          if ((s.count1 > 1) && (hrat > 2.0)) {
            s.hnew = s.hg;
            break;
          }
          s.hg = s.hnew;
        }
      }
      double h0 = 0.5 * s.hnew;
      if (h0 < s.hlb) {
        h0 = s.hlb;
      }
      if (h0 > s.hub) {
        h0 = s.hub;
      }
      s.h = h0;
    }
    s.hscale = s.h;
    s.hprime = s.h;
    PRX_VEC_FOR(i) {
      s.zn(1)[i] *= s.h;
    }
    s.nstloc = 0;
    // ---- main loop of CVode (CV_NORMAL, no tstop) ----
    for (;;) {
      if (s.nst > 0) {
        if (ewt_set(s, p, s.zn(0)) != 0) {
          s.flag = CV_ILL_INPUT;
          break;
        }
      }
      if ((p.mxstep > 0) && (s.nstloc >= p.mxstep)) {
        s.flag = CV_TOO_MUCH_WORK;
        break;
      }
      if (UROUND * wrms(s, s.zn(0), s.ewt()) > 1.0) {
        s.flag = CV_TOO_MUCH_ACC;
        break;
      }
      // ---- cvStep ----
      s.ncf = s.nef = 0;
      if ((s.nst > 0) && (s.hprime != s.h)) {
        if (s.qprime != s.q) {
          bdf_adjust_order(s, s.qprime - s.q, p.qmax);
          s.q = s.qprime;
          s.L = s.q + 1;
          s.qwait = s.L;
        }
        bdf_rescale(s);
      }
      s.saved_t = s.tn;
      s.nflag = FIRST_CALL;
      for (;;) { // attempts at this step
        bdf_predict(s);
        bdf_set(s);
        // ---- cvNls ----
        s.convfail = ((s.nflag == FIRST_CALL) || (s.nflag == PREV_ERR_FAIL)) ? 0 : 2;
        s.callSetup = (s.nflag == PREV_CONV_FAIL) || (s.nflag == PREV_ERR_FAIL) || (s.nst == 0) ||
                      (s.nst >= s.nstlp + 20) || (fabs(s.gamrat - 1.0) > 0.3);
        s.jbad_nls = 0;
        s.nls_jcur = 0;
        s.nls_ok = 0;
        PRX_VEC_FOR(i) {
          s.acor()[i] = 0.0;
        }
        for (;;) { // Newton attempts (second pass: with a fresh Jacobian)
          s.curiter = 0;
          PRX_VEC_FOR(i) {
            s.y()[i] = s.zn(0)[i] + s.acor()[i];
          }
          PRX_YIELD_RHS(3, s.tn, V_Y, V_FTEMP);
          s.nfe++;
          PRX_VEC_FOR(i) {
            double d = s.rl1 * s.zn(1)[i] + s.acor()[i];
            d = -s.gamma * s.ftemp()[i] + d;
            s.delta()[i] = d;
          }
          s.retval = 0;
          if (s.callSetup) {
            if (s.jbad_nls) {
              s.convfail = 1;
            }
            // ---- cvLsSetup ----
            {
              const double dgamma = fabs((s.gamma / s.gammap) - 1.0);
              const int jbad = (s.nst == 0) || (s.nst >= s.nstlj + p.msbj) ||
                               ((s.convfail == 1) && (dgamma < 0.2)) || (s.convfail == 2);
              s.jcur = jbad;
            }
            if (s.jcur) {
              if (p.analytic_jac) {
                PRX_YIELD(4, REQ_JAC); // kernel fills savedJ with J(tn, y)
              } else {
                // cvLsDenseDQJac
                {
                  const double fnorm = wrms(s, s.ftemp(), s.ewt());
                  s.minInc = (fnorm != 0.0) ? (1000.0 * fabs(s.h) * UROUND * NEQ * fnorm) : 1.0;
                }
                for (s.jcol = 0; s.jcol < NEQ; s.jcol++) {
                  s.sync();
                  s.yjsaved = s.y()[s.jcol];
                  s.inc = fmax(sqrt(UROUND) * fabs(s.yjsaved), s.minInc / s.ewt()[s.jcol]);
                  s.sync();
                  if (s.owns(s.jcol)) {
                    s.y()[s.jcol] += s.inc;
                  }
                  PRX_YIELD_RHS(5, s.tn, V_Y, V_TEMPV);
                  s.nfeDQ++;
                  if (s.owns(s.jcol)) {
                    s.y()[s.jcol] = s.yjsaved;
                  }
                  {
                    const double inc_inv = 1.0 / s.inc;
                    PRX_VEC_FOR(i) {
                      s.savedJ()[s.jcol * NEQ + i] = inc_inv * (s.tempv()[i] - s.ftemp()[i]);
                    }
                  }
                }
                // the RHS may use delta as scratch (shared-memory kernels): rebuild the Newton residual,
                // same operands and operations as above, hence the same bits
                PRX_VEC_FOR(i) {
                  double d = s.rl1 * s.zn(1)[i] + s.acor()[i];
                  d = -s.gamma * s.ftemp()[i] + d;
                  s.delta()[i] = d;
                }
              }
              s.nje++;
              s.nstlj = s.nst;
            }
            // A = I - gamma*J from the saved Jacobian, in one pass (the reference copies savedJ -> A,
            // scales by -gamma and adds the identity: the same two roundings per entry)
            {
              const double mg = -s.gamma;
              s.sync();
              if (s.vstep() == 1) { // one thread per cell: column by column, the diagonal test is compile-time
                PRX_ALOOP
                for (int j = 0; j < NEQ; j++) {
                  PRX_VEC_FOR(i) {
                    double v = s.savedJ()[j * NEQ + i];
                    v *= mg;
                    if (i == j) {
                      v += 1.0;
                    }
                    s.A()[j * NEQ + i] = v;
                  }
                }
              } else {
                PRX_ALOOP
                for (int e = s.vbegin(); e < NEQ * NEQ; e += s.vstep()) {
                  double v = s.savedJ()[e];
                  v *= mg;
                  if (e % (NEQ + 1) == 0) { // diagonal of the column-major matrix
                    v += 1.0;
                  }
                  s.A()[e] = v;
                }
              }
            }
            PRX_YIELD(6, REQ_LU); // kernel factors A in place, sets piv and lu_info
            s.nsetups++;
            s.nls_jcur = s.jcur;
            s.gamrat = 1.0;
            s.gammap = s.gamma;
            s.crate = 1.0;
            s.nstlp = s.nst;
            if (s.lu_info > 0) {
              s.retval = NLS_CONV_RECVR;
            }
          }
          if (s.retval == 0) {
            for (;;) { // Newton iterations
              s.nni++;
              PRX_VEC_FOR(i) {
                s.delta()[i] = -s.delta()[i];
              }
              s.getrs(s.A(), s.piv(), s.delta());
              if (s.gamrat != 1.0) {
                const double sc = 2.0 / (1.0 + s.gamrat);
                PRX_VEC_FOR(i) {
                  s.delta()[i] *= sc;
                }
              }
              PRX_VEC_FOR(i) {
                s.acor()[i] = s.acor()[i] + s.delta()[i];
              }
              {
                const double del = wrms(s, s.delta(), s.ewt());
                const int m = s.curiter;
                if (m > 0) {
                  s.crate = fmax(0.3 * s.crate, del / s.delp);
                }
                const double dcon = del * fmin(1.0, s.crate) / s.tq[4];
                if (dcon <= 1.0) {
                  s.acnrm = (m == 0) ? del : wrms(s, s.acor(), s.ewt());
                  s.nls_ok = 1;
                  s.retval = 0;
                } else if ((m >= 1) && (del > 2.0 * s.delp)) {
                  s.retval = NLS_CONV_RECVR;
                } else {
                  s.delp = del;
                  s.retval = NLS_CONTINUE;
                }
              }
              s.curiter++;
              if (s.retval != NLS_CONTINUE) {
                break;
              }
              if (s.curiter >= p.maxcor) {
                s.retval = NLS_CONV_RECVR;
                break;
              }
              PRX_VEC_FOR(i) {
                s.y()[i] = s.zn(0)[i] + s.acor()[i];
              }
              PRX_YIELD_RHS(7, s.tn, V_Y, V_FTEMP);
              s.nfe++;
              PRX_VEC_FOR(i) {
                double d = s.rl1 * s.zn(1)[i] + s.acor()[i];
                d = -s.gamma * s.ftemp()[i] + d;
                s.delta()[i] = d;
              }
            }
          }
          if (s.nls_ok) {
            break;
          }
          if ((s.retval > 0) && !s.nls_jcur) {
            s.callSetup = 1;
            s.jbad_nls = 1;
            PRX_VEC_FOR(i) {
              s.acor()[i] = 0.0;
            }
            continue;
          }
          break;
        }
        if (s.nls_ok) {
          s.jcur = 0;
          PRX_VEC_FOR(i) {
            s.y()[i] = s.zn(0)[i] + s.acor()[i];
          }
        } else {
          // ---- cvHandleNFlag: recoverable convergence failure ----
          s.ncfn++;
          bdf_restore(s);
          s.ncf++;
          s.etamax = 1.0;
          if (s.ncf == 10) {
            s.flag = CV_CONV_FAILURE;
            break;
          }
          s.eta = 0.25;
          bdf_rescale(s);
          s.nflag = PREV_CONV_FAIL;
          continue;
        }
        // ---- cvDoErrorTest ----
        s.dsm = s.acnrm * s.tq[2];
        if (s.dsm <= 1.0) {
          break;
        }
        s.nef++;
        s.netf++;
        s.nflag = PREV_ERR_FAIL;
        bdf_restore(s);
        if (s.nef == p.maxnef) {
          s.flag = CV_ERR_FAILURE;
          break;
        }
        s.etamax = 1.0;
        if (s.nef <= 3) {
          s.eta = 1.0 / (rpowerR(6.0 * s.dsm, 1.0 / s.L) + 0.000001);
          s.eta = fmax(0.1, s.eta);
          if (s.nef >= 2) {
            s.eta = fmin(s.eta, 0.2);
          }
          bdf_rescale(s);
          continue;
        }
        if (s.q > 1) {
          s.eta = 0.1;
          bdf_adjust_order(s, -1, p.qmax);
          s.L = s.q;
          s.q--;
          s.qwait = s.L;
          bdf_rescale(s);
          continue;
        }
        s.eta = 0.1;
        s.h *= s.eta;
        s.hscale = s.h;
        s.qwait = 10;
        PRX_YIELD_RHS(8, s.tn, V_ZN0, V_TEMPV);
        s.nfe++;
        PRX_VEC_FOR(i) {
          s.zn(1)[i] = s.h * s.tempv()[i];
        }
      }
      if (s.flag != CV_SUCCESS) {
        break;
      }
      // ---- cvCompleteStep ----
      s.nst++;
      s.hu = s.h;
      s.qu = s.q;
      for (int i = s.q; i >= 2; i--) {
        s.tau[i] = s.tau[i - 1];
      }
      if ((s.q == 1) && (s.nst > 1)) {
        s.tau[2] = s.tau[1];
      }
      s.tau[1] = s.h;
      PRX_QLOOP
      for (int j = 0; j < LMAX; j++) { // j <= q, written with a constant trip count so that it can be unrolled
        if (j <= s.q) {
          const double lj = s.l[j];
          PRX_VEC_FOR(i) {
            s.zn(j)[i] += lj * s.acor()[i];
          }
        }
      }
      s.qwait--;
      if ((s.qwait == 1) && (s.q != p.qmax)) {
        PRX_VEC_FOR(i) {
          s.zn(p.qmax)[i] = s.acor()[i];
        }
        s.saved_tq5 = s.tq[5];
        s.indx_acor = p.qmax;
      }
      bdf_prepare_next_step(s, p.qmax);
      s.etamax = 10.0;
      {
        const double t2 = s.tq[2];
        PRX_VEC_FOR(i) {
          s.acor()[i] *= t2;
        }
      }
      s.nstloc++;
      if ((s.tn - s.tout) * s.h >= 0.0) {
        // CVodeGetDky(tout, 0): Horner-free sum in the reference's N_VLinearCombination order
        const double sfrac = (s.tout - s.tn) / s.h;
        double cvals[LMAX];
        int nvec = 0;
        for (int j = s.q; j >= 0; j--) {
          double c = 1.0;
          for (int i = 0; i < j; i++) {
            c *= sfrac;
          }
          cvals[nvec++] = c;
        }
        PRX_VEC_FOR(i) {
          double z = cvals[0] * s.zn(s.q)[i];
          for (int j = 1; j < nvec; j++) {
            z += cvals[j] * s.zn(s.q - j)[i];
          }
          s.y()[i] = z;
        }
        s.pc = 99;
        return REQ_DONE;
      }
    }
    // failure exits: the reference ignores CVode's flag on the CPU path and keeps yout = zn[0]
    PRX_VEC_FOR(i) {
      s.y()[i] = s.zn(0)[i];
    }
    break;
  default:
    break;
  }
  s.pc = 99;
  return REQ_DONE;
}

} // namespace prx
#endif
