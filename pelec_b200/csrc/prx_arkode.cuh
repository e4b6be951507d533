// prx_arkode.cuh - ReactorArkode (PP/Source/Reactions/ReactorArkode.cpp:135-260, ode.use_erkstep = 1) behind the C ABI.
// Included by prx_api.cu after prx_handle is defined.
//
// What the reference does, on every backend: the whole box is ONE ODE system of neq * ncells unknowns handed to ARKODE's
// ERKStep module - embedded explicit Runge-Kutta (default Zonneveld 5-3-4), one step size for all cells, error test on the
// WRMS norm over all cells, PID step-size controller, Hermite interpolation back to tout.  Here:
//   * every stage is ONE kernel over the cells (ops->ark_stage: stage state + right-hand side, per mechanism) instead of the
//     reference's N_VLinearCombination + RHS launch pair; the solution / error-estimate combination and its norm are one
//     kernel too;
//   * the step control is this file's host code, a PORT of the scalar logic of ARKODE for this configuration - no root
//     finding, no tstop, no constraints, no relaxation, no fixed steps:
//       ARKodeEvolve            SUN/src/arkode/arkode.c:636-1090
//       arkInitialSetup         arkode.c:1789-2075     (initial error weights, h0, first-step growth factor)
//       arkHin / arkYddNorm / arkUpperBoundH0   arkode.c:2295-2483
//       arkCheckTemporalError   arkode.c:3113-3170,  arkAdapt  arkode_adapt.c:95-185, with ERKStep's overrides of the
//                               adaptivity defaults (safety 0.99, growth 25, etamxf 0.3: arkode_erkstep_io.c:310-312)
//       arkCompleteStep         arkode.c:2489-2570
//       erkStep_TakeStep / erkStep_ComputeSolutions / erkStep_FullRHS   arkode_erkstep.c:502-1060
//       SUNAdaptController Soderlind (PID / PI / I)   SUN/src/sunadaptcontroller/soderlind/sunadaptcontroller_soderlind.c:308-425
//       arkInterpUpdate/Evaluate_Hermite (cubic)      arkode_interp.c:381-560,  ARKodeGetDky  arkode.c:1230-1290
//   * one reference quirk is kept: PelePhysics creates the SUNAdaptController once in init() and attaches it to every new
//     ERKStep object without a reset (ReactorArkode.cpp:90-116, 207; ARKodeSetAdaptController does not reset), so the
//     controller's error / step history carries over from one react call to the next.  It lives in the handle here.
// Norms are summed per block and then over the blocks in a fixed order (deterministic, but not the left-to-right order of
// N_VWrmsNorm_Serial): step sizes agree with the reference to rounding, results within the integrator's tolerance.
#ifndef PRX_ARKODE_CUH
#define PRX_ARKODE_CUH

namespace prx {
namespace ark {

struct Table
{
  int stages, q, p;
  double A[5][5], b[5], d[5], c[5];
};

// Butcher tables, SUN/src/arkode/arkode_butcher_erk.def (method, embedding)
inline bool table_for(int method, Table* T)
{
  std::memset(T, 0, sizeof(*T));
  switch (method) {
  case 20: // ARKODE_HEUN_EULER_2_1_2
    T->stages = 2; T->q = 2; T->p = 1;
    T->A[1][0] = 1.0;
    T->b[0] = 0.5; T->b[1] = 0.5;
    T->d[0] = 1.0;
    T->c[1] = 1.0;
    return true;
  case 30: // ARKODE_BOGACKI_SHAMPINE_4_2_3
    T->stages = 4; T->q = 3; T->p = 2;
    T->A[1][0] = 1.0 / 2.0;
    T->A[2][1] = 3.0 / 4.0;
    T->A[3][0] = 2.0 / 9.0; T->A[3][1] = 1.0 / 3.0; T->A[3][2] = 4.0 / 9.0;
    T->b[0] = 2.0 / 9.0; T->b[1] = 1.0 / 3.0; T->b[2] = 4.0 / 9.0;
    T->d[0] = 7.0 / 24.0; T->d[1] = 1.0 / 4.0; T->d[2] = 1.0 / 3.0; T->d[3] = 1.0 / 8.0;
    T->c[1] = 1.0 / 2.0; T->c[2] = 3.0 / 4.0; T->c[3] = 1.0;
    return true;
  case 40: // ARKODE_ZONNEVELD_5_3_4
    T->stages = 5; T->q = 4; T->p = 3;
    T->A[1][0] = 0.5;
    T->A[2][1] = 0.5;
    T->A[3][2] = 1.0;
    T->A[4][0] = 5.0 / 32.0; T->A[4][1] = 7.0 / 32.0; T->A[4][2] = 13.0 / 32.0; T->A[4][3] = -1.0 / 32.0;
    T->b[0] = 1.0 / 6.0; T->b[1] = 1.0 / 3.0; T->b[2] = 1.0 / 3.0; T->b[3] = 1.0 / 6.0;
    T->d[0] = -1.0 / 2.0; T->d[1] = 7.0 / 3.0; T->d[2] = 7.0 / 3.0; T->d[3] = 13.0 / 6.0; T->d[4] = -16.0 / 3.0;
    T->c[1] = 0.5; T->c[2] = 0.5; T->c[3] = 1.0; T->c[4] = 0.75;
    return true;
  default:
    return false;
  }
}

// ---- mechanism-independent vector kernels (CYOrder columns, n = neq * ncells entries) ---------------------------------
constexpr int RB = 256;  // threads per block
constexpr int RG = 592;  // blocks (4 per SM): partial results, reduced in a fixed order by finish_kernel

__device__ __forceinline__ double block_sum(double v)
{
  __shared__ double sh[RB / 32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    v += __shfl_xor_sync(0xffffffffu, v, o);
  }
  if ((threadIdx.x & 31) == 0) {
    sh[threadIdx.x >> 5] = v;
  }
  __syncthreads();
  double r = 0.0;
  if (threadIdx.x == 0) {
    for (int w = 0; w < RB / 32; w++) {
      r += sh[w];
    }
  }
  __syncthreads();
  return r;
}
__device__ __forceinline__ double block_max(double v)
{
  __shared__ double shm[RB / 32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  }
  if ((threadIdx.x & 31) == 0) {
    shm[threadIdx.x >> 5] = v;
  }
  __syncthreads();
  double r = 0.0;
  if (threadIdx.x == 0) {
    for (int w = 0; w < RB / 32; w++) {
      r = fmax(r, shm[w]);
    }
  }
  __syncthreads();
  return r;
}

// out[0] = sum of part[0..RG), out[1] = max of part[RG..2RG)
__global__ void finish_kernel(const double* __restrict__ part, double* __restrict__ out)
{
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    double s = 0.0, m = 0.0;
    for (int b = 0; b < RG; b++) {
      s += part[b];
      m = fmax(m, part[RG + b]);
    }
    out[0] = s;
    out[1] = m;
  }
}

// arkEwtSetSV (arkode.c:2734-2760): ewt = 1 / (rtol |y| + atol_k); partial sums of (y ewt)^2 (the "too much accuracy" test)
// and the largest rtol|y| + atol <= 0 flag (as max of a 0/1 indicator)
__global__ void ewt_kernel(
  const long long n, const long long nc, const double* __restrict__ y, const double rtol, const double* __restrict__ atolv,
  double* __restrict__ ewt, double* __restrict__ part)
{
  double s = 0.0, bad = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const double w = rtol * fabs(y[i]) + atolv[i / nc];
    if (!(w > 0.0)) {
      bad = 1.0;
    }
    const double e = 1.0 / w;
    ewt[i] = e;
    const double p = y[i] * e;
    s += p * p;
  }
  const double bs = block_sum(s), bm = block_max(bad);
  if (threadIdx.x == 0) {
    part[blockIdx.x] = bs;
    part[RG + blockIdx.x] = bm;
  }
}

// arkUpperBoundH0 (arkode.c:2415-2445): max |fn| / (0.1 |yn| + 1 / ewt)
__global__ void hub_kernel(
  const long long n, const double* __restrict__ yn, const double* __restrict__ fn, const double* __restrict__ ewt,
  double* __restrict__ part)
{
  double m = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const double t1 = 0.1 * fabs(yn[i]) + 1.0 * (1.0 / ewt[i]);
    m = fmax(m, fabs(fn[i]) / t1);
  }
  const double bm = block_max(m);
  if (threadIdx.x == 0) {
    part[blockIdx.x] = 0.0;
    part[RG + blockIdx.x] = bm;
  }
}

// ycur = hg fn + yn (arkYddNorm, first half)
__global__ void axpy_kernel(
  const long long n, const double a, const double* __restrict__ x, const double* __restrict__ y, double* __restrict__ z)
{
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    z[i] = a * x[i] + y[i];
  }
}

// arkYddNorm, second half: ydd = (1/hg) f1 + (-1/hg) fn, partial sums of (ydd ewt)^2
__global__ void ydd_kernel(
  const long long n, const double ihg, const double* __restrict__ f1, const double* __restrict__ fn,
  const double* __restrict__ ewt, double* __restrict__ part)
{
  double s = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const double v = (ihg * f1[i] + (-ihg) * fn[i]) * ewt[i];
    s += v * v;
  }
  const double bs = block_sum(s);
  if (threadIdx.x == 0) {
    part[blockIdx.x] = bs;
    part[RG + blockIdx.x] = 0.0;
  }
}

// erkStep_ComputeSolutions: ycur = sum_j (h b_j) F_j + yn; yerr = sum_j h (b_j - d_j) F_j; partial sums of (yerr ewt)^2
struct CombineArgs
{
  int stages;
  double cb[5], ce[5];
  const double* F[5];
};
__global__ void combine_kernel(
  const long long n, const CombineArgs a, const double* __restrict__ yn, const double* __restrict__ ewt,
  double* __restrict__ ycur, double* __restrict__ part)
{
  double s = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    double y = a.cb[0] * a.F[0][i], e = a.ce[0] * a.F[0][i];
    for (int j = 1; j < a.stages; j++) {
      const double f = a.F[j][i];
      y += a.cb[j] * f;
      e += a.ce[j] * f;
    }
    y += 1.0 * yn[i];
    ycur[i] = y;
    const double v = e * ewt[i];
    s += v * v;
  }
  const double bs = block_sum(s);
  if (threadIdx.x == 0) {
    part[blockIdx.x] = bs;
    part[RG + blockIdx.x] = 0.0;
  }
}

// Hermite interpolant of degree 1 / 2 / 3 (arkode_interp.c cases 1-3, d = 0): yout = a0 x0 + a1 x1 [+ a2 x2 [+ a3 x3]] in
// N_VLinearCombination's order; unused terms are passed with coefficient 0 and a valid pointer
__global__ void interp_kernel(
  const long long n, const double a0, const double a1, const double a2, const double a3, const double* __restrict__ yold,
  const double* __restrict__ yn, const double* __restrict__ fold, const double* __restrict__ fn, double* __restrict__ yout)
{
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    double v = a0 * yold[i];
    v += a1 * yn[i];
    if (a2 != 0.0) {
      v += a2 * fold[i];
    }
    if (a3 != 0.0) {
      v += a3 * fn[i];
    }
    yout[i] = v;
  }
}

__global__ void fill_ll_kernel(const long long n, const long long v, long long* __restrict__ p)
{
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    p[i] = v;
  }
}

// SUNAdaptController (Soderlind family): the state PelePhysics keeps across react calls
struct Controller
{
  double k1, k2, k3, bias;
  double ep, epp, hp, hpp;
  int firststeps;
  void init(int kind)
  {
    k1 = 0.58; k2 = -0.21; k3 = 0.1; // PID
    if (kind == 1) { k1 = 0.8; k2 = -0.31; k3 = 0.0; } // PI
    if (kind == 2) { k1 = 1.0; k2 = 0.0; k3 = 0.0; }   // I
    bias = 1.5;
    ep = epp = hp = hpp = 1.0;
    firststeps = 0;
  }
  static double powr(double base, double ex) { return base <= 0.0 ? 0.0 : std::pow(base, ex); } // SUNRpowerR
  double estimate(double h, int p, double dsm) const
  {
    const int ord = p + 1;
    const double e1 = std::fmax(bias * dsm, 1.0e-10), e2 = std::fmax(ep, 1.0e-10), e3 = std::fmax(epp, 1.0e-10);
    const double q1 = -k1 / ord, q2 = -k2 / ord, q3 = -k3 / ord;
    const double hrat = h / hp, hrat2 = hp / hpp; // exponents k4 = k5 = 0 for PID / PI / I: pow(x, 0) = 1
    if (firststeps < 1) {
      return h * powr(e1, q1);
    }
    if (firststeps < 2) {
      return h * powr(e1, q1) * powr(e2, q2) * powr(hrat, 0.0);
    }
    return h * powr(e1, q1) * powr(e2, q2) * powr(e3, q3) * powr(hrat, 0.0) * powr(hrat2, 0.0);
  }
  void update(double h, double dsm)
  {
    epp = ep;
    ep = bias * dsm;
    hpp = hp;
    hp = h;
    if (firststeps < 2) {
      firststeps += 1;
    }
  }
};

} // namespace ark
} // namespace prx
#endif
