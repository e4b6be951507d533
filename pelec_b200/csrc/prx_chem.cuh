// prx_chem.cuh - per-cell thermo-chemistry on the device: the EOS facade and the ODE
// right-hand side, layered on the generated mechanism header (gen/mech_<mech>.cuh).
//
// Reference functions restated here (PP = PelePhysics):
//   GET_T_GIVEN_EY / GET_T_GIVEN_HY   PP/Mechanisms/LiDryer/mechanism.H:1419-1520
//   CKUMS / CKHMS / CKYTCR / CKWC     mechanism.H:2019-2043, 1785-1796, 2606-2623
//   Fuego::RTY2WDOT / RTY2JAC         PP/Source/Eos/Fuego.H:234-266 (+DWDOT mechanism.H:5141-5162)
//   utils::fKernelSpec                PP/Source/Reactions/ReactorUtils.H:358-415
//   cvode::cJac (CPU flavour)         PP/Source/Reactions/ReactorCvodeJacobian.cpp:86-152
//
// One translation unit per mechanism (PRX_MECH_HEADER selects it), like the reference's
// compile-time Chemistry_Model.  Everything is FP64; all per-cell vectors live in registers /
// thread-local storage of the calling kernel.
#ifndef PRX_CHEM_CUH
#define PRX_CHEM_CUH

#ifdef PRX_HOST_TEST
// Host instantiation used ONLY by tests/host_harness (logic unit tests without a GPU).
#include <cmath>
#define PRX_DEVFN static inline
#define PRX_DEVCONST static const
#define PRX_DEV static inline
#define PRX_DEV_NOINLINE static
#define PRX_MEM inline
#define PRX_HD inline
#else
#define PRX_DEVFN __device__ __forceinline__
#define PRX_DEVCONST static __constant__ const
#define PRX_DEV __device__ __forceinline__
#define PRX_DEV_NOINLINE __device__ __noinline__
#define PRX_MEM __device__ __forceinline__
#define PRX_HD __host__ __device__ __forceinline__
#endif

#if !defined(PRX_HOST_TEST)
// Transcendentals of the generated straight-line code (gen.py redirects exp / log / log10 here).
// ncu, round 1 (profiles/r01n_ncu_bdf_lidryer_64_by_line.txt): libm's exp was 27 % and log10 8 % of ALL executed
// instructions of the BDF kernel - 44 instructions per exp of which 26 are UMOV pairs for its 13 literals, 146 per log10.
//  * prx_exp (PRX_CUSTOM_MATH=1): the same algorithm and coefficients as the library's fast path (argument reduction with a
//    two-part ln 2, degree-11 polynomial, exponent patched into the high word; 0.90 ulp max on [-708, 708] measured against
//    expl on the host), literals from a __constant__ table (one LDCU.128 per two).  Arguments outside the fast range (the
//    reference's exp(-T*1e30) Troe terms, overflow, NaN) take the library path.
//  * prx_log10 = log(x) * log10(e): <= 1.5 ulp instead of the library's 1 ulp, 41 instructions instead of 146.
// Measured (session r02d, 128^3 LiDryer): BDF 66.4 vs 66.7 ms, RK64 92.7 vs 90.9 ms - no gain: a single out-of-line exp is a
// chain of 13 dependent DFMAs whose latency, not its instruction count, is what the warp waits for.  OFF by default (the
// library functions keep the round-1 parity numbers); kept for the batched-evaluation experiment.
#ifndef PRX_CUSTOM_MATH
#define PRX_CUSTOM_MATH 0
#endif
#if defined(PRX_NOINLINE_MATH)
#define PRX_MATHFN static __device__ __noinline__
#else
#define PRX_MATHFN static __device__ __forceinline__
#endif
static __constant__ __align__(16) double prx_mathk[16] = {
  1.4426950408889634, 6755399441055744.0, -6.9314718055994529e-01, -2.3190468138462996e-17,
  2.5052097064908941e-08, 2.7626262793835868e-07, 2.7557414788000726e-06, 2.4801504602132958e-05,
  1.9841269707468915e-04, 1.3888888932258898e-03, 8.3333333333978320e-03, 4.1666666666573905e-02,
  1.6666666666666563e-01, 5.0000000000000056e-01, 0.43429448190325182, 0.0};
#if PRX_CUSTOM_MATH
PRX_MATHFN double prx_exp(const double x)
{
  if (!(fabs(x) <= 708.0)) {
    return exp(x);
  }
  double t = fma(x, prx_mathk[0], prx_mathk[1]);
  const int i = __double2loint(t);
  t -= prx_mathk[1];
  double a = fma(t, prx_mathk[2], x);
  a = fma(t, prx_mathk[3], a);
  double q = prx_mathk[4];
  q = fma(q, a, prx_mathk[5]);
  q = fma(q, a, prx_mathk[6]);
  q = fma(q, a, prx_mathk[7]);
  q = fma(q, a, prx_mathk[8]);
  q = fma(q, a, prx_mathk[9]);
  q = fma(q, a, prx_mathk[10]);
  q = fma(q, a, prx_mathk[11]);
  q = fma(q, a, prx_mathk[12]);
  q = fma(q, a, prx_mathk[13]);
  q = fma(q, a, 1.0);
  q = fma(q, a, 1.0);
  return __hiloint2double(__double2hiint(q) + (i << 20), __double2loint(q));
}
PRX_MATHFN double prx_log(const double x) { return log(x); }
PRX_MATHFN double prx_log10(const double x) { return prx_log(x) * prx_mathk[14]; }
#else
PRX_MATHFN double prx_exp(const double x) { return exp(x); }
PRX_MATHFN double prx_log(const double x) { return log(x); }
PRX_MATHFN double prx_log10(const double x) { return log10(x); }
#endif
#define PRX_MATH_REDIRECT 1
#endif

// ---- paired exponentials (gen.py emits PRX_EXP2(a, b) for the two independent exp() of a reversible reaction) -----------------
struct prx_d2
{
  double x, y;
};
#ifndef PRX_BATCH_EXP
#define PRX_BATCH_EXP 1
#endif
#if !defined(PRX_HOST_TEST) && PRX_BATCH_EXP
// Both chains of the library's fast-path algorithm (see prx_exp above: same reduction, same degree-11 polynomial, same
// coefficients) advance in lock step, so that every DFMA has an independent neighbour to hide its latency behind; literals
// are materialised once for the two values.  Arguments outside [-708, 708] / NaN: library path for both.
static __device__ __noinline__ prx_d2 prx_exp2(const double xa, const double xb)
{
  prx_d2 r;
  if (!(fabs(xa) <= 708.0) || !(fabs(xb) <= 708.0)) {
    r.x = exp(xa);
    r.y = exp(xb);
    return r;
  }
  double ta = fma(xa, 1.4426950408889634, 6755399441055744.0);
  double tb = fma(xb, 1.4426950408889634, 6755399441055744.0);
  const int ia = __double2loint(ta), ib = __double2loint(tb);
  ta -= 6755399441055744.0;
  tb -= 6755399441055744.0;
  double a = fma(ta, -6.9314718055994529e-01, xa), b = fma(tb, -6.9314718055994529e-01, xb);
  a = fma(ta, -2.3190468138462996e-17, a);
  b = fma(tb, -2.3190468138462996e-17, b);
  double p = 2.5052097064908941e-08, q = 2.5052097064908941e-08;
#define PRX_E2(c)      \
  p = fma(p, a, (c)); \
  q = fma(q, b, (c));
  PRX_E2(2.7626262793835868e-07)
  PRX_E2(2.7557414788000726e-06)
  PRX_E2(2.4801504602132958e-05)
  PRX_E2(1.9841269707468915e-04)
  PRX_E2(1.3888888932258898e-03)
  PRX_E2(8.3333333333978320e-03)
  PRX_E2(4.1666666666573905e-02)
  PRX_E2(1.6666666666666563e-01)
  PRX_E2(5.0000000000000056e-01)
  PRX_E2(1.0)
  PRX_E2(1.0)
#undef PRX_E2
  r.x = __hiloint2double(__double2hiint(p) + (ia << 20), __double2loint(p));
  r.y = __hiloint2double(__double2hiint(q) + (ib << 20), __double2loint(q));
  return r;
}
#define PRX_EXP2(a, b) prx_exp2((a), (b))
#else
#define PRX_EXP2(a, b) (prx_d2{exp(a), exp(b)})
#endif

#include PRX_MECH_HEADER

#define PRX_N PRX_NUM_SPECIES
#define PRX_NEQ (PRX_NUM_SPECIES + 1)

namespace prx {

constexpr int N = PRX_N;
constexpr int NEQ = PRX_NEQ;

// Newton solve of e(T,Y) = e (energy_type 1) or h(T,Y) = h (2): identical control flow to the
// reference (bounds 90/4000 K with linear extrapolation, +-100 K clamp, break BEFORE the update).
template <int ETYPE>
PRX_DEV void t_from_energy(const double e, const double* y, double& t)
{
  const double tmin = PRX_TMIN, tmax = PRX_TMAX;
  const double emin = (ETYPE == 1) ? ckubms(tmin, y) : ckhbms(tmin, y);
  const double emax = (ETYPE == 1) ? ckubms(tmax, y) : ckhbms(tmax, y);
  if (e < emin) {
    const double cv = (ETYPE == 1) ? ckcvbs(tmin, y) : ckcpbs(tmin, y);
    t = tmin - (emin - e) / cv;
    return;
  }
  if (e > emax) {
    const double cv = (ETYPE == 1) ? ckcvbs(tmax, y) : ckcpbs(tmax, y);
    t = tmax - (emax - e) / cv;
    return;
  }
  double t1 = t;
  if (t1 < tmin || t1 > tmax) {
    t1 = tmin + (tmax - tmin) / (emax - emin) * (e - emin);
  }
  for (int i = 0; i < 200; ++i) {
    const double e1 = (ETYPE == 1) ? ckubms(t1, y) : ckhbms(t1, y);
    const double cv = (ETYPE == 1) ? ckcvbs(t1, y) : ckcpbs(t1, y);
    double dt = (e - e1) / cv;
    if (dt > 100.) {
      dt = 100.;
    } else if (dt < -100.) {
      dt = -100.;
    } else if (fabs(dt) < 1.e-6) {
      break;
    }
    t1 += dt;
  }
  t = t1;
}

PRX_DEV void t_from_energy(const int reactor_type, const double e, const double* y, double& t)
{
  if (reactor_type == 1) {
    t_from_energy<1>(e, y, t);
  } else {
    t_from_energy<2>(e, y, t);
  }
}

// Per-cell constants of the ODE (CVODEUserData rhoe_init / rhoesrc_ext / rYsrc_ext,
// PP/Source/Reactions/ReactorCvodeUtils.H:34-100)
struct CellSrc
{
  double rhoe_init;
  double rhoesrc;
  double rYsrc[N];
};

// The same constants with the species source as a COLUMN of the block's shared memory (element k at p[k * STRIDE],
// STRIDE = threads per block: conflict-free).  The RHS reads rYsrc at its very end, 40 times per cell, from thread-local
// memory that has left L1 by then (ncu r02i: 6.6 % of the gen-1 BDF kernel's stall samples sit on that line); the first
// generation's blocks use no shared memory otherwise.
template <int STRIDE>
struct CellSrcCol
{
  double rhoe_init;
  double rhoesrc;
  struct Col
  {
    double* p;
    PRX_MEM double& operator[](const int i) const { return p[i * STRIDE]; }
  } rYsrc;
};

// ODE right-hand side for one cell: y = (rhoY_0..rhoY_{N-1}, T) -> ydot.
template <class SrcT>
PRX_DEV void rhs(const double t, const int reactor_type, const double* y, double* ydot, const SrcT& c)
{
  double rho = 0.0;
  double Y[N];
#pragma unroll
  for (int n = 0; n < N; n++) {
    rho += y[n];
  }
  const double rho_inv = 1.0 / rho;
#pragma unroll
  for (int n = 0; n < N; n++) {
    Y[n] = y[n] * rho_inv;
  }
  const double nrg = (c.rhoe_init + c.rhoesrc * t) * rho_inv;
  double temp = y[N];
  double Cv;
  double ei[N];
  const double RT_unused = 0.0;
  (void)RT_unused;
  if (reactor_type == 1) {
    t_from_energy<1>(nrg, Y, temp);
    e_RT_fn(ei, temp);
    Cv = ckcvbs(temp, Y);
  } else {
    t_from_energy<2>(nrg, Y, temp);
    h_RT_fn(ei, temp);
    Cv = ckcpbs(temp, Y);
  }
  const double RT = PRX_RU * temp;
#pragma unroll
  for (int i = 0; i < N; i++) {
    ei[i] *= RT * prx_imw[i];
  }
  // RTY2WDOT: C = rho*Y/W (CGS) -> SI -> production_rate -> CGS -> mass rates
  double C[N], wdot[N];
#pragma unroll
  for (int i = 0; i < N; i++) {
    C[i] = rho * Y[i] * prx_imw[i];
    C[i] *= 1.0e6;
  }
  production_rate(wdot, C, temp);
  double rhoesrc = c.rhoesrc;
#pragma unroll
  for (int n = 0; n < N; n++) {
    double w = wdot[n] * 1.0e-6;
    w *= prx_mw[n];
    const double cdot_rYs = w + c.rYsrc[n];
    ydot[n] = cdot_rYs;
    rhoesrc -= cdot_rYs * ei[n];
  }
  ydot[N] = rhoesrc * (rho_inv / Cv);
}

// Analytic Jacobian of the ODE in (rhoY, T) variables, written through the accessor
// JM(row, col) (= d f_row / d y_col).  Scratch Jt is NEQ*NEQ doubles, column-major.
template <class Store>
PRX_DEV void jac_analytic(const int reactor_type, const double* ydata, double* Jt, Store&& JM)
{
  double rho = 0.0;
#pragma unroll
  for (int i = 0; i < N; i++) {
    rho = rho + ydata[i];
  }
  const double temp = ydata[N];
  double c[N];
#pragma unroll
  for (int i = 0; i < N; i++) {
    const double Yi = ydata[i] / rho;
    c[i] = 1.e6 * (rho * Yi * prx_imw[i]);
  }
  const int consP = (reactor_type == 2);
  ajacobian(Jt, c, temp, consP);
  // DWDOT unit conversion (SI <-> CGS) and the cJac scaling between C and rhoY variables
  for (int i = 0; i < N; i++) {
    for (int k = 0; k < N; k++) {
      JM(k, i, Jt[i * NEQ + k] * prx_mw[k] * prx_imw[i]);
    }
    JM(N, i, (Jt[i * NEQ + N] * 1.e6) * prx_imw[i]);
  }
  for (int i = 0; i < N; i++) {
    JM(i, N, (Jt[N * NEQ + i] * 1.e-6) * prx_mw[i]);
  }
  JM(N, N, Jt[N * NEQ + N]);
}

} // namespace prx
#endif
