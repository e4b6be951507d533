// prx_kernels.cu - the sm_100a kernels of the reactor; compiled ONCE PER MECHANISM with
//   -DPRX_MECH_HEADER="gen/mech_<mech>.cuh" -DPRX_MECH_SUFFIX=<mech>
//
// Kernel anatomy (both reactors): a PERSISTENT grid (a multiple of the 148 SMs) whose warps
// loop over three phases until the box's work queue is empty:
//   (1) refill   lanes without a cell take the next cell indices from a global atomic queue
//                (warp-aggregated: one atomicAdd per warp per refill).  Cells are handed out
//                in index order, so a warp's loads/stores of the SoA arrays stay mostly
//                coalesced along i and neighbouring (similarly stiff) cells share a warp.
//                This is the active-cell compaction: a lane never idles because its cell
//                finished early, it just starts the next one.
//   (2) control  every lane runs its integrator coroutine (prx_integrators.cuh) up to the
//                next expensive request - cheap, divergent FP64 vector code.
//   (3) serve    the warp executes each kind of request (RHS / Jacobian / LU) once, at a
//                converged call site, for all lanes that asked.  Rare kinds can be deferred
//                until several lanes ask (SchedParams), which does not change any cell's
//                arithmetic, only when it happens.
// HBM traffic is only the boundary arrays: (2N+3) doubles in, (N+3) out per cell.
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "prx_integrators.cuh"
#include "prx_tchem.cuh"
#include "prx_coop.cuh"
#include "prx_kernels.cuh"

#ifndef PRX_RK_BLOCK // block shape of the RK64 kernel (its register appetite differs from the BDF kernel's)
#define PRX_RK_BLOCK PRX_BLOCK
#endif
#ifndef PRX_RK_MINBLOCKS
#define PRX_RK_MINBLOCKS PRX_MINBLOCKS
#endif
#ifndef PRX_BLOCK
#define PRX_BLOCK 128
#endif

#ifndef PRX_MINBLOCKS
#define PRX_MINBLOCKS 1
#endif

#ifndef PRX_INLINE_SLOW
#define PRX_INLINE_SLOW 0
#endif
#ifndef PRX_PREFETCH_SRC
#define PRX_PREFETCH_SRC 0
#endif
#ifndef PRX_PREFETCH_STATE
#define PRX_PREFETCH_STATE 0
#endif
#ifndef PRX_BDF_DEFAULT_VARIANT
#define PRX_BDF_DEFAULT_VARIANT 0
#endif

#define PRX_CAT2(a, b) a##b
#define PRX_CAT(a, b) PRX_CAT2(a, b)

namespace prx {
namespace {

constexpr unsigned FULL = 0xffffffffu;

struct LaneStats
{
  unsigned long long cells, masked, steps, rhs, jac, setups, newton, errfail, convfail, failed, maxsteps, maxrhs;
};

__device__ __forceinline__ unsigned long long warp_sum(unsigned long long v)
{
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    v += __shfl_xor_sync(FULL, v, o);
  }
  return v;
}
__device__ __forceinline__ unsigned long long warp_max(unsigned long long v)
{
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long w = __shfl_xor_sync(FULL, v, o);
    v = w > v ? w : v;
  }
  return v;
}

__device__ __forceinline__ void flush_stats(const LaneStats& ls, DevStats* st)
{
  const unsigned lane = threadIdx.x & 31;
  unsigned long long v[11] = {ls.cells, ls.masked, ls.steps, ls.rhs, ls.jac, ls.setups,
                              ls.newton, ls.errfail, ls.convfail, ls.failed, 0ull};
#pragma unroll
  for (int i = 0; i < 10; i++) {
    v[i] = warp_sum(v[i]);
  }
  const unsigned long long mx = warp_max(ls.maxsteps);
  const unsigned long long mr = warp_max(ls.maxrhs);
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < 10; i++) {
      if (v[i]) {
        atomicAdd(&st->v[i], v[i]);
      }
    }
    atomicMax(&st->v[ST_MAXSTEPS], mx);
    atomicAdd(&st->v[ST_LAST_STEPS], v[2]);
    atomicMax(&st->v[ST_LAST_MAXRHS], mr);
  }
}

// warp-aggregated fetch of the next cell indices; returns -1 when the queue is empty
__device__ __forceinline__ int fetch_cell(const bool need, int* queue, const int ncells)
{
  const unsigned lane = threadIdx.x & 31;
  const unsigned m = __ballot_sync(FULL, need);
  if (m == 0) {
    return -1;
  }
  int base = 0;
  const int leader = __ffs(m) - 1;
  if (lane == leader) {
    base = atomicAdd(queue, __popc(m));
  }
  base = __shfl_sync(FULL, base, leader);
  if (!need) {
    return -1;
  }
  const int idx = base + __popc(m & ((1u << lane) - 1u));
  return (idx < ncells) ? idx : -2; // -2: exhausted
}

__device__ __forceinline__ long long cell_offset(const int c, const int* len, const long long js, const long long ks)
{
  const int i = c % len[0];
  const int r = c / len[0];
  const int j = r % len[1];
  const int k = r / len[1];
  return (long long)i + (long long)j * js + (long long)k * ks;
}

// ---- fused PeleC::react_state body (PC/Source/React.cpp:181-343), one cell ---------------------------------
// Pack (React.cpp:181-220): the reactor integrates from the OLD state (rhoY, T, rho*e_int of S_old) with the non-reacting
// species sources and the energy source rhoedot_ext = (rho*e_int implied by S_new = S_old + dt*sources  -  rho_old*e_old) / dt.
// Rotating frame (pelec.do_rf): rotenrg = get_rot_energy (PC/Source/Utilities.H:23-65) of the cell centre; 0 when off, and then
// every term below that carries it is exact-zero arithmetic, as in the reference.
__device__ __forceinline__ double state_rotenrg(const BoxArgs& a, const int c)
{
  if (!a.rf_on) {
    return 0.0;
  }
  const int i = c % a.len[0];
  const int r = c / a.len[0];
  const int idx[3] = {a.lo[0] + i, a.lo[1] + r % a.len[1], a.lo[2] + r / a.len[1]};
  double rv[3];
#pragma unroll
  for (int d = 0; d < 3; d++) {
    rv[d] = (a.rf_problo[d] + (idx[d] + 0.5) * a.rf_dx[d]) - a.rf_axloc[d]; // pc_cmp_loc - axis location
  }
  double rad2 = rv[0] * rv[0] + rv[1] * rv[1] + rv[2] * rv[2]; // RealVect::radSquared
  rad2 -= rv[a.rf_axis] * rv[a.rf_axis];                         // only in plane
  const double rad = sqrt(rad2);
  return 0.5 * a.rf_omega * a.rf_omega * rad * rad;
}

__device__ __forceinline__ double state_e_old(const BoxArgs& a, const long long oo, double& rho_old, const double rotenrg)
{
  const double* so = a.sold.p + oo;
  const long long ns = a.sold.ns;
  const double rhou = so[a.ix.umx * ns], rhov = so[a.ix.umy * ns], rhow = so[a.ix.umz * ns];
  rho_old = so[a.ix.urho * ns];
  const double rhoInv = 1.0 / rho_old;
  double e_old = (so[a.ix.ueden * ns] - 0.5 * (rhou * rhou + rhov * rhov + rhow * rhow) * rhoInv) * rhoInv;
  e_old += rotenrg;
  return e_old;
}

__device__ __forceinline__ double state_rhoedot_ext(const BoxArgs& a, const int c)
{
  double rho_old;
  const double rotenrg = state_rotenrg(a, c);
  const double e_old = state_e_old(a, cell_offset(c, a.len, a.sold.js, a.sold.ks), rho_old, rotenrg);
  const double* sn = a.snew.p + cell_offset(c, a.len, a.snew.js, a.snew.ks);
  const long long ns = a.snew.ns;
  const double rhou = sn[a.ix.umx * ns], rhov = sn[a.ix.umy * ns], rhow = sn[a.ix.umz * ns];
  const double rhoInv = 1.0 / sn[a.ix.urho * ns];
  return (sn[a.ix.ueden * ns] - 0.5 * (rhou * rhou + rhov * rhov + rhow * rhow) * rhoInv + sn[a.ix.urho * ns] * rotenrg -
          rho_old * e_old) /
         a.dt;
}

// Unpack + I_R + heat release (React.cpp:233-343) from the reactor's output (rhoY_new[N], T_new) of one cell.
__device__ __forceinline__ void state_unpack(
  const BoxArgs& a, const int c, const double* rhoY, const double Tnew, const double rhoedot_ext, const double fc)
{
  const long long oo = cell_offset(c, a.len, a.sold.js, a.sold.ks);
  const long long on = cell_offset(c, a.len, a.snew.js, a.snew.ks);
  const long long os = cell_offset(c, a.len, a.nonrs.js, a.nonrs.ks);
  const long long oi = cell_offset(c, a.len, a.IR.js, a.IR.ks);
  const double* so = a.sold.p + oo;
  const double* nr = a.nonrs.p + os;
  double* sn = a.snew.p + on;
  double* ir = a.IR.p + oi;
  const long long sos = a.sold.ns, nrs = a.nonrs.ns, sns = a.snew.ns, irs = a.IR.ns;
  const double dt = a.dt;
  double rho_old;
  const double rotenrg = state_rotenrg(a, c);
  const double e_old = state_e_old(a, oo, rho_old, rotenrg);
  const double umnew = so[a.ix.umx * sos] + dt * nr[a.ix.umx * nrs];
  const double vmnew = so[a.ix.umy * sos] + dt * nr[a.ix.umy * nrs];
  const double wmnew = so[a.ix.umz * sos] + dt * nr[a.ix.umz * nrs];
  double rhonew = 0.0;
#pragma unroll
  for (int n = 0; n < N; n++) {
    rhonew += rhoY[n];
  }
  const double eint_new = rho_old * e_old + dt * rhoedot_ext;
  const double eden_new = eint_new + 0.5 * (umnew * umnew + vmnew * vmnew + wmnew * wmnew) / rhonew - rhonew * rotenrg;
  double T_h = Tnew;
  if (a.do_update) {
    sn[a.ix.urho * sns] = rhonew;
    sn[a.ix.umx * sns] = umnew;
    sn[a.ix.umy * sns] = vmnew;
    sn[a.ix.umz * sns] = wmnew;
#pragma unroll
    for (int n = 0; n < N; n++) {
      sn[(a.ix.ufs + n) * sns] = rhoY[n];
    }
    sn[a.ix.utemp * sns] = Tnew;
    sn[a.ix.ueint * sns] = eint_new;
    sn[a.ix.ueden * sns] = eden_new;
  } else {
    T_h = sn[a.ix.utemp * sns]; // react_init: S_new is not touched, the heat release is evaluated at its temperature
  }
  double hi[N];
  h_RT_fn(hi, T_h); // eos.RTY2Hi -> CKHMS (mechanism.H:2032-2041)
  const double RT = PRX_RU * T_h;
  double hr = 0.0;
#pragma unroll
  for (int n = 0; n < N; n++) {
    const double irn = (rhoY[n] - so[(a.ix.ufs + n) * sos]) / dt - nr[(a.ix.ufs + n) * nrs];
    ir[n * irs] = irn;
    hr -= (hi[n] * (RT * prx_imw[n])) * irn;
  }
  ir[N * irs] = (eden_new - so[a.ix.ueden * sos]) / dt - nr[a.ix.ueden * nrs];
  ir[(N + 1) * irs] = hr;
  if (a.FC.p != nullptr) {
    a.FC.p[cell_offset(c, a.len, a.FC.js, a.FC.ks)] = fc;
  }
}

// ---- boundary loads / stores ---------------------------------------------------------
// Array4: SoA, consecutive lanes <-> consecutive i: every component load is a coalesced
// 256-byte warp transaction.  Flat (YCOrder): cell-major, used by the 1-D react() only.
template <class SrcT>
__device__ __forceinline__ bool load_cell(
  const BoxArgs& a, const int c, double* y /*NEQ*/, SrcT& src, const int clean_init_massfrac)
{
  if (a.layout == LAYOUT_ARRAY4) {
    if (a.mask.p != nullptr) {
      if (a.mask.p[cell_offset(c, a.len, a.mask.js, a.mask.ks)] == -1) {
        return false;
      }
    }
    const long long o = cell_offset(c, a.len, a.rY.js, a.rY.ks);
    const long long os = cell_offset(c, a.len, a.rYsrc.js, a.rYsrc.ks);
    double rho_old = 0.0;
#pragma unroll
    for (int k = 0; k < N; k++) {
      y[k] = __ldg(a.rY.p + o + k * a.rY.ns);
      src.rYsrc[k] = __ldg(a.rYsrc.p + os + k * a.rYsrc.ns);
      rho_old += y[k];
    }
    if (clean_init_massfrac) { // ReactorUtils.H:183-199
#pragma unroll
      for (int k = 0; k < N; k++) {
        y[k] = fmin(fmax(0.0, y[k]), rho_old);
      }
    }
    y[N] = a.T.p[cell_offset(c, a.len, a.T.js, a.T.ks)];
    src.rhoe_init = a.rE.p[cell_offset(c, a.len, a.rE.js, a.rE.ks)];
    src.rhoesrc = a.rEsrc.p[cell_offset(c, a.len, a.rEsrc.js, a.rEsrc.ks)];
  } else if (a.layout == LAYOUT_STATE) {
    const double* so = a.sold.p + cell_offset(c, a.len, a.sold.js, a.sold.ks);
    const double* nr = a.nonrs.p + cell_offset(c, a.len, a.nonrs.js, a.nonrs.ks);
    double rho_old = 0.0;
#pragma unroll
    for (int k = 0; k < N; k++) {
      y[k] = __ldg(so + (a.ix.ufs + k) * a.sold.ns);
      src.rYsrc[k] = __ldg(nr + (a.ix.ufs + k) * a.nonrs.ns);
      rho_old += y[k];
    }
    if (clean_init_massfrac) {
#pragma unroll
      for (int k = 0; k < N; k++) {
        y[k] = fmin(fmax(0.0, y[k]), rho_old);
      }
    }
    y[N] = so[a.ix.utemp * a.sold.ns];
    src.rhoe_init = so[a.ix.ueint * a.sold.ns];
    src.rhoesrc = state_rhoedot_ext(a, c);
  } else {
#pragma unroll
    for (int k = 0; k < NEQ; k++) {
      y[k] = a.y[(long long)c * NEQ + k];
    }
#pragma unroll
    for (int k = 0; k < N; k++) {
      src.rYsrc[k] = a.ysrc[(long long)c * N + k];
    }
    src.rhoe_init = a.e[c];
    src.rhoesrc = a.esrc[c];
  }
  return true;
}

__device__ __forceinline__ void store_masked(const BoxArgs& a, const int c)
{
  a.FC.p[cell_offset(c, a.len, a.FC.js, a.FC.ks)] = 0.0;
}

// box_unflatten (ReactorUtils.H:222-291) / RK64 write-back (ReactorRK64.cpp:290-309)
template <class SrcT>
__device__ __forceinline__ void store_cell(
  const BoxArgs& a, const int c, const double* y /*NEQ: rhoY_new, T (pre re-solve)*/, const SrcT& src,
  const int reactor_type, const double fc, const int clean_init_massfrac, const bool resolve_T)
{
  const double rhoe_out = src.rhoe_init + a.dt * src.rhoesrc;
  if (a.layout == LAYOUT_STATE) {
    // box_unflatten semantics on the reactor's side (rhoY out, T re-solved from (e_out, Y_out)), then React.cpp's unpack
    double ynew[N];
    double rho = 0.0;
    const double* so = a.sold.p + cell_offset(c, a.len, a.sold.js, a.sold.ks);
    if (clean_init_massfrac) {
      double rho_old = 0.0, rho_oldclipped = 0.0;
#pragma unroll
      for (int k = 0; k < N; k++) {
        rho_old += so[(a.ix.ufs + k) * a.sold.ns];
      }
#pragma unroll
      for (int k = 0; k < N; k++) {
        rho_oldclipped += fmin(fmax(0.0, so[(a.ix.ufs + k) * a.sold.ns]), rho_old);
      }
#pragma unroll
      for (int k = 0; k < N; k++) {
        const double yold = so[(a.ix.ufs + k) * a.sold.ns];
        ynew[k] = yold + (y[k] - fmin(fmax(0.0, yold), rho_old)) * rho_old / rho_oldclipped;
      }
    } else {
#pragma unroll
      for (int k = 0; k < N; k++) {
        ynew[k] = y[k];
      }
    }
#pragma unroll
    for (int k = 0; k < N; k++) {
      rho += ynew[k];
    }
    double temp = y[N];
    if (resolve_T) {
      const double rho_inv = 1.0 / rho;
      double Y[N];
#pragma unroll
      for (int k = 0; k < N; k++) {
        Y[k] = ynew[k] * rho_inv;
      }
      t_from_energy(reactor_type, rhoe_out * rho_inv, Y, temp);
    }
    state_unpack(a, c, ynew, temp, src.rhoesrc, fc);
    return;
  }
  if (a.layout == LAYOUT_ARRAY4) {
    const long long o = cell_offset(c, a.len, a.rY.js, a.rY.ks);
    double ynew[N];
    double rho = 0.0;
    if (clean_init_massfrac) {
      // new rhoY = old + CVODE increment measured against the clipped input (ReactorUtils.H:244-271)
      double rho_old = 0.0, rho_oldclipped = 0.0;
      double yold[N];
#pragma unroll
      for (int k = 0; k < N; k++) {
        yold[k] = a.rY.p[o + k * a.rY.ns];
        rho_old += yold[k];
      }
#pragma unroll
      for (int k = 0; k < N; k++) {
        rho_oldclipped += fmin(fmax(0.0, yold[k]), rho_old);
      }
#pragma unroll
      for (int k = 0; k < N; k++) {
        const double clipped = fmin(fmax(0.0, yold[k]), rho_old);
        ynew[k] = yold[k] + (y[k] - clipped) * rho_old / rho_oldclipped;
      }
    } else {
#pragma unroll
      for (int k = 0; k < N; k++) {
        ynew[k] = y[k];
      }
    }
#pragma unroll
    for (int k = 0; k < N; k++) {
      a.rY.p[o + k * a.rY.ns] = ynew[k];
      rho += ynew[k];
    }
    double temp = y[N];
    if (resolve_T) {
      const double rho_inv = 1.0 / rho;
      double Y[N];
#pragma unroll
      for (int k = 0; k < N; k++) {
        Y[k] = ynew[k] * rho_inv;
      }
      t_from_energy(reactor_type, rhoe_out * rho_inv, Y, temp);
    }
    a.rE.p[cell_offset(c, a.len, a.rE.js, a.rE.ks)] = rhoe_out;
    a.T.p[cell_offset(c, a.len, a.T.js, a.T.ks)] = temp;
    a.FC.p[cell_offset(c, a.len, a.FC.js, a.FC.ks)] = fc;
  } else {
#pragma unroll
    for (int k = 0; k < NEQ; k++) {
      a.y[(long long)c * NEQ + k] = y[k];
    }
    a.e[c] = rhoe_out;
    if (a.fc != nullptr) {
      a.fc[c] = fc;
    }
  }
}

// =====================================================================================
// RK64 kernel
// =====================================================================================
__global__ void __launch_bounds__(PRX_RK_BLOCK, PRX_RK_MINBLOCKS) PRX_CAT(rk64_kernel_, PRX_MECH_SUFFIX)(
  const BoxArgs a, const Rk64Params p, int* __restrict__ queue, DevStats* __restrict__ stats)
{
  Rk64Cell s;
  int cell = -1;
  bool exhausted = false;
  int req = REQ_NONE;
  LaneStats ls = {};
  for (;;) {
    int got = fetch_cell((cell < 0) && !exhausted, queue, a.ncells);
    if (got >= 0 && a.order != nullptr) {
      got = a.order[got]; // cells handed out in sorted order (see prx_api.cu build_order)
    }
    if (got == -2) {
      exhausted = true;
    } else if (got >= 0) {
      cell = got;
      load_cell(a, cell, s.soln, s.src, 0); // RK64 ignores the mask (reference: /*mask*/ unused)
      rk64_begin(s, p, a.dt, a.time);
      req = REQ_NONE;
    }
    // block-wide exit test; it doubles as the phase barrier that keeps the block's warps walking
    // through the (large, straight-line) RHS code together, so that one instruction-cache fill
    // serves every warp of the SM (ncu: stall_no_instruction was 76 % of all samples without it)
    if (!__syncthreads_or(!(exhausted && (cell < 0)))) {
      break;
    }
    if (cell >= 0 && req == REQ_NONE) {
      req = rk64_resume(s, p);
    }
    if (req == REQ_RHS) {
      rhs(s.t_eval, p.reactor_type, s.soln, s.ydot, s.src);
      req = REQ_NONE;
    }
    if (req == REQ_DONE) {
      // the write-back re-solves T from (e_out, Y_out) with the integrated T as the guess
      store_cell(a, cell, s.soln, s.src, p.reactor_type, double(s.nsteps), 0, p.array4_semantics != 0);
      ls.cells++;
      ls.steps += s.nsteps;
      ls.rhs += 6ull * s.nsteps;
      ls.maxsteps = (unsigned long long)s.nsteps > ls.maxsteps ? s.nsteps : ls.maxsteps;
      cell = -1;
      req = REQ_NONE;
    }
  }
  flush_stats(ls, stats);
}

// =====================================================================================
// BDF (CVODE-per-cell) kernel
// =====================================================================================
__device__ __forceinline__ double* bdf_vec(BdfCell& s, const int sel)
{
  switch (sel) {
  case V_ZN0:
    return s.zn(0);
  case V_ZN1:
    return s.zn(1);
  case V_Y:
    return s.y();
  case V_TEMPV:
    return s.tempv();
  default:
    return s.ftemp();
  }
}

// one cell per thread with the straight-line generated Jacobian only makes sense for small mechanisms (and the
// 16k-line dodecane_lu Jacobian costs six minutes of compile time): large ones use the cooperative kernel
#ifdef PRX_AUDIT_RK64_ONLY // audit build (-fmad=false, Makefile target `audit`): RK64 + probes only
#define PRX_HAVE_BDF1 0
#else
#define PRX_HAVE_BDF1 (PRX_NUM_SPECIES <= 32)
#endif
#ifndef PRX_SMEM_SRC
#define PRX_SMEM_SRC 0
#endif
extern __shared__ double prx_dyn_smem[];
constexpr int PRX_BDF1_SMEM_BYTES =
  PRX_SMEM_SRC ? (PRX_NUM_SPECIES + ((PRX_SMEM_SRC >= 2) ? PRX_NUM_SPECIES + 1 : 0) + ((PRX_SMEM_SRC >= 3) ? PRX_NUM_SPECIES + 1 : 0)) * PRX_BLOCK * 8 : 0;
#if PRX_HAVE_BDF1
#ifndef PRX_OUTLINE
#define PRX_OUTLINE 0 // bit 0: Jacobian out of line, bit 1: RHS out of line (separate register allocation, fewer spills in the loop)
#endif
#if PRX_OUTLINE & 1
__device__ __noinline__ void jac_outline(const int reactor_type, const double* y, double* J)
{
  jac_analytic(reactor_type, y, J, [J](const int row, const int col, const double v) { J[col * NEQ + row] = v; });
}
#endif
#if PRX_OUTLINE & 2
__device__ __noinline__ void rhs_outline(const double t, const int reactor_type, const double* y, double* ydot, const CellSrc* c)
{
  rhs(t, reactor_type, y, ydot, *c);
}
#endif
__global__ void __launch_bounds__(PRX_BLOCK, PRX_MINBLOCKS) PRX_CAT(bdf_kernel_, PRX_MECH_SUFFIX)(
  const BoxArgs a, const BdfParams p, const SchedParams sp, int* __restrict__ queue,
  DevStats* __restrict__ stats)
{
#if PRX_SMEM_SRC
  // the cell's external species source (and, PRX_SMEM_SRC >= 2 / 3, ewt / acor) in shared-memory columns (CellSrcCol,
  // BdfStoreLocalHot): [rYsrc N][ewt NEQ][acor NEQ] x PRX_BLOCK doubles of dynamic shared memory
  BdfCellT<BdfStoreLocalHot<PRX_BLOCK, PRX_SMEM_SRC>, CellSrcCol<PRX_BLOCK>> s;
  s.src.rYsrc.p = prx_dyn_smem + threadIdx.x;
  s.hot = prx_dyn_smem + N * PRX_BLOCK + threadIdx.x;
#else
  BdfCell s;
#endif
  int cell = -1;
  bool exhausted = false;
  int req = REQ_NONE;
  int slow_age = 0;
  LaneStats ls = {};
  for (;;) {
    // ---- (1) refill ----
    // Refill in batches: with unrelated neighbouring cells a lane that restarts alone evaluates its first Jacobian and
    // factorisation alone (1/32 lane efficiency); holding idle lanes back until refill_min of them are free makes the
    // new cells start - and ask for their Jacobians - together.
    bool want = (cell < 0) && !exhausted;
    if (sp.refill_min > 1) {
      const unsigned m_want = __ballot_sync(FULL, want);
      const unsigned m_busy = __ballot_sync(FULL, cell >= 0);
      if (m_busy != 0u && __popc(m_want) < sp.refill_min) {
        want = false;
      }
    }
    int got = fetch_cell(want, queue, a.ncells);
    if (got >= 0 && a.order != nullptr) {
      got = a.order[got]; // cells handed out in sorted order (see prx_api.cu build_order)
    }
    if (got == -2) {
      exhausted = true;
    } else if (got >= 0) {
      if (load_cell(a, got, s.zn(0), s.src, p.clean_init_massfrac)) {
        cell = got;
        bdf_begin(s, p, a.dt, a.time);
        req = REQ_NONE;
      } else {
        store_masked(a, got); // mask == -1: FC = 0, nothing else touched (ReactorCvode.cpp:1451-1453)
        ls.masked++;
      }
    }
    // exit test; with sp.block_warps != 1 it is block wide and doubles as the phase barrier (see rk64 kernel);
    // mode 2 has a single barrier per iteration, the one in front of the RHS, which then carries the exit test
    if (sp.block_warps == 2) {
    } else if (sp.block_warps != 1) {
      if (!__syncthreads_or(!(exhausted && (cell < 0)))) {
        break;
      }
    } else if (!__any_sync(FULL, !(exhausted && (cell < 0)))) {
      break;
    }
    // ---- (2) control ----
#if PRX_INLINE_SLOW
    // Jacobian / LU requests are served inside the control phase (warp-level loop), so that every iteration of
    // the block ends in exactly one RHS evaluation per warp and no iteration is spent on one warp's factorisation
    for (;;) {
      if (cell >= 0 && req == REQ_NONE) {
        req = bdf_resume(s, p);
      }
      if (__ballot_sync(FULL, req == REQ_JAC || req == REQ_LU) == 0u) {
        break;
      }
      if (req == REQ_JAC) {
        double* J = s.savedJ();
#if PRX_OUTLINE & 1
        jac_outline(p.reactor_type, s.y(), J);
#else
        jac_analytic(p.reactor_type, s.y(), J, [J](const int row, const int col, const double v) {
          J[col * NEQ + row] = v;
        });
#endif
        req = REQ_NONE;
      }
      if (req == REQ_LU) {
        s.lu_info = (NEQ <= 12) ? dense_getrf_small(s.A(), s.piv(), s.perm()) : ((NEQ <= 24) ? dense_getrf_mid(s.A(), s.piv(), s.perm()) : dense_getrf(s.A(), s.piv()));
        req = REQ_NONE;
      }
    }
#else
    if (cell >= 0 && req == REQ_NONE) {
      req = bdf_resume(s, p);
    }
#endif
    // ---- (3) serve ----
    // Block-level batching of the slow requests (Jacobian, LU): with the phase barrier every iteration costs
    // as much as its slowest warp, and with 8 warps almost every iteration had ONE warp factoring (ncu: 32 % of
    // the stall samples on this barrier).  A warp with a slow request therefore waits (for at most
    // lu_defer_max_age iterations) until lu_defer_block percent of the block's threads have one too, or until
    // nobody in the block has RHS work; then they all factor in the same iteration.
    bool slow_ok = true;
    if (sp.block_warps == 2) {
      if (!__syncthreads_or(!(exhausted && (cell < 0)))) {
        break;
      }
    } else if (sp.block_warps != 1 && sp.block_warps != 3) {
      if (sp.lu_defer_block > 0) {
        const int n_slow = __syncthreads_count(req == REQ_JAC || req == REQ_LU);
        const int n_fast = __syncthreads_count(req == REQ_RHS);
        slow_ok = (n_slow * 100 >= sp.lu_defer_block * (int)blockDim.x) || (n_fast == 0) || (slow_age >= sp.lu_defer_max_age);
        if (req == REQ_JAC || req == REQ_LU) {
          slow_age = slow_ok ? 0 : slow_age + 1;
        }
      } else {
        __syncthreads(); // re-align the warps after the divergent control code, right before the RHS
      }
    }
    const unsigned m_rhs = __ballot_sync(FULL, req == REQ_RHS);
    const unsigned m_jac = __ballot_sync(FULL, req == REQ_JAC && slow_ok);
    const unsigned m_lu = __ballot_sync(FULL, req == REQ_LU && slow_ok);
    if (req == REQ_RHS) {
#if PRX_PREFETCH_SRC
      // the external species source is read at the very end of the RHS, from local memory that has usually left L1
      // by then (ncu r01k: 6 % of the stall samples on that line): start the fetch now
#pragma unroll
      for (int k = 0; k < N; k++) {
        asm volatile("prefetch.L1 [%0];" ::"l"(&s.src.rYsrc[k]));
      }
#endif
#if PRX_PREFETCH_STATE
      // The control phase that follows the RHS reads the Newton matrix (triangular solves) and the hot vectors from local
      // memory that has left L1 / L2 while 20 other warps ran (ncu r02i: L1 hit 41 %, L2 hit 51 %, long-scoreboard 41 % of the
      // stall samples): start those fetches now, behind ~5 k instructions of register-only chemistry.  MEASURED (r02y, 128^3
      // LiDryer): 67.4 ms without, 71.7 ms with the matrix prefetched (L1 or L2), 74.1 ms with matrix + vectors - the kernel's
      // memory system is saturated by the state traffic itself, extra requests only queue.  OFF.  Local memory interleaves
      // the lanes of a warp by 32-bit word, so a double is two 128-byte rows.  Bit 0: the matrix, bit 1: zn0, zn1, acor, ewt,
      // tempv; bit 2: into L1 instead of L2.
#define PRX_PF(ptr)                                                                         \
  do {                                                                                      \
    if (PRX_PREFETCH_STATE & 4) {                                                           \
      asm volatile("prefetch.L1 [%0];" ::"l"(ptr));                                         \
      asm volatile("prefetch.L1 [%0];" ::"l"(reinterpret_cast<const char*>(ptr) + 4));      \
    } else {                                                                                \
      asm volatile("prefetch.L2 [%0];" ::"l"(ptr));                                         \
      asm volatile("prefetch.L2 [%0];" ::"l"(reinterpret_cast<const char*>(ptr) + 4));      \
    }                                                                                       \
  } while (0)
      if (PRX_PREFETCH_STATE & 1) {
#pragma unroll
        for (int k = 0; k < NEQ * NEQ; k++) {
          PRX_PF(&s.A_[k]);
        }
      }
      if (PRX_PREFETCH_STATE & 2) {
#pragma unroll
        for (int k = 0; k < NEQ; k++) {
          PRX_PF(&s.zn_[0][k]);
          PRX_PF(&s.zn_[1][k]);
          PRX_PF(&s.acor_[k]);
          PRX_PF(&s.ewt_[k]);
          PRX_PF(&s.tempv_[k]);
        }
      }
#undef PRX_PF
#endif
#if PRX_OUTLINE & 2
      rhs_outline(s.t_eval, p.reactor_type, s.vec(s.in_sel), s.vec(s.out_sel), &s.src);
#else
      rhs(s.t_eval, p.reactor_type, s.vec(s.in_sel), s.vec(s.out_sel), s.src);
#endif
      req = REQ_NONE;
    }
#if !PRX_INLINE_SLOW
    if (m_jac && ((__popc(m_jac) >= sp.jac_defer_min_lanes) || (m_rhs == 0))) {
      if (req == REQ_JAC && slow_ok) {
        double* J = s.savedJ(); // raw aJacobian output, scaled in place (each entry is read once, then written)
        jac_analytic(p.reactor_type, s.y(), J, [J](const int row, const int col, const double v) {
          J[col * NEQ + row] = v;
        });
        req = REQ_NONE;
      }
    }
    if (m_lu && ((__popc(m_lu) >= sp.lu_defer_min_lanes) || (m_rhs == 0))) {
      if (req == REQ_LU && slow_ok) {
        s.lu_info = (NEQ <= 12) ? dense_getrf_small(s.A(), s.piv(), s.perm()) : ((NEQ <= 24) ? dense_getrf_mid(s.A(), s.piv(), s.perm()) : dense_getrf(s.A(), s.piv()));
        req = REQ_NONE;
      }
    }
#endif
    if (req == REQ_DONE) {
      store_cell(
        a, cell, s.y(), s.src, p.reactor_type, double(s.nfe + s.nfeDQ), p.clean_init_massfrac,
        p.array4_semantics != 0);
      ls.cells++;
      ls.steps += s.nst;
      ls.rhs += s.nfe + s.nfeDQ;
      ls.maxrhs = (unsigned long long)(s.nfe + s.nfeDQ) > ls.maxrhs ? (unsigned long long)(s.nfe + s.nfeDQ) : ls.maxrhs;
      ls.jac += s.nje;
      ls.setups += s.nsetups;
      ls.newton += s.nni;
      ls.errfail += s.netf;
      ls.convfail += s.ncfn;
      ls.failed += (s.flag != CV_SUCCESS);
      ls.maxsteps = (unsigned long long)s.nst > ls.maxsteps ? s.nst : ls.maxsteps;
      cell = -1;
      req = REQ_NONE;
    }
  }
  flush_stats(ls, stats);
}
#endif // PRX_HAVE_BDF1

// =====================================================================================
// BDF kernel, second generation: integrator state in SHARED-MEMORY COLUMNS + table-driven RHS
// =====================================================================================
// What the first kernel's ncu capture showed (profiles/r01a_ncu_bdf_lidryer_64.csv): 478 KB of SASS,
// 76 % of the stall samples "no instruction", 4.3 KB of thread-local stack per cell spilling to DRAM
// (28 KB of DRAM traffic per cell against 268 B algorithmic).  Here
//  * every per-cell vector (zn[0..qmax], ewt, acor, tempv, ftemp, y, delta, a species scratch, and -
//    template flag - the Newton matrix A) is a column of dynamic shared memory: element i of a vector
//    sits at base[i*32 + lane], so a warp access is one conflict-free 256-byte row and the state never
//    leaves the SM;
//  * the RHS is the rolled, table-driven evaluation of prx_tchem.cuh (mechanism tables in constant
//    memory: the 32 lanes of a warp evaluate the same reaction, so each table read is a broadcast);
//    the hot loop (control + RHS + LU + solves) is a few tens of KB of SASS;
//  * blocks are single warps: no block barrier anywhere, the SM scheduler interleaves independent
//    warps (as many as the shared memory holds: launch_bdf computes it with the occupancy API);
//  * the external species source is not copied per cell, it is re-read (coalesced, L2-resident)
//    through SrcRef at every RHS.
struct SrcRef
{
  double rhoe_init, rhoesrc;
  const double* p; // rYsrc of this cell, component k at p[k*stride]
  long long stride;
};


#ifndef PRX_FULL_VARIANTS
#define PRX_FULL_VARIANTS 0 // 1: also build the measured dead ends (bdf2 / bdf3 / bdf4, profiles/README.md) for A/B runs
#endif
#define PRX_HAVE_BDF2 (PRX_FULL_VARIANTS && PRX_NUM_SPECIES <= 32) // one cell per thread stops making sense beyond that
#if PRX_HAVE_BDF2

template <bool A_SMEM>
__global__ void __launch_bounds__(32) PRX_CAT(bdf2_kernel_, PRX_MECH_SUFFIX)(
  const BoxArgs a, const BdfParams p, const SchedParams sp, int* __restrict__ queue, DevStats* __restrict__ stats)
{
  typedef BdfCellT<BdfStoreSmem<32, A_SMEM>, SrcRef> Cell;
  Cell s;
  s.base = prx_dyn_smem + threadIdx.x;
  s.nz = p.qmax + 1;
  int cell = -1;
  bool exhausted = false;
  int req = REQ_NONE;
  LaneStats ls = {};
  BdfParams pc = p;
  pc.array4_semantics = 0; // the initial T solve is done here with the table-driven Newton loop
  for (;;) {
    // ---- (1) refill ----
    int got = fetch_cell((cell < 0) && !exhausted, queue, a.ncells);
    if (got >= 0 && a.order != nullptr) {
      got = a.order[got]; // cells handed out in sorted order (see prx_api.cu build_order)
    }
    if (got == -2) {
      exhausted = true;
    } else if (got >= 0) {
      double y0[NEQ];
      CellSrc cs;
      if (load_cell(a, got, y0, cs, p.clean_init_massfrac)) {
        cell = got;
        s.src.rhoe_init = cs.rhoe_init;
        s.src.rhoesrc = cs.rhoesrc;
        if (a.layout == LAYOUT_ARRAY4) {
          s.src.p = a.rYsrc.p + cell_offset(got, a.len, a.rYsrc.js, a.rYsrc.ks);
          s.src.stride = a.rYsrc.ns;
        } else {
          s.src.p = a.ysrc + (long long)got * N;
          s.src.stride = 1;
        }
        double rho = 0.0;
#pragma unroll
        for (int k = 0; k < N; k++) {
          s.zn(0)[k] = y0[k];
          rho += y0[k];
        }
        double temp = y0[N];
        if (p.array4_semantics) { // box_flatten re-solves T from (e, Y) (ReactorUtils.H:206-219)
          const double rho_inv = 1.0 / rho;
#pragma unroll
          for (int k = 0; k < N; k++) {
            s.scratch()[k] = y0[k] * rho_inv;
          }
          t_from_energy_tab<TabC>(p.reactor_type, cs.rhoe_init * rho_inv, s.scratch(), temp);
        }
        s.zn(0)[N] = temp;
        bdf_begin(s, pc, a.dt, a.time);
        req = REQ_NONE;
      } else {
        store_masked(a, got);
        ls.masked++;
      }
    }
    if (!__any_sync(FULL, !(exhausted && (cell < 0)))) {
      break;
    }
    // ---- (2) control ----
    if (cell >= 0 && req == REQ_NONE) {
      req = bdf_resume(s, pc);
    }
    __syncwarp();
    // ---- (3) serve ----
    const unsigned m_rhs = __ballot_sync(FULL, req == REQ_RHS);
    const unsigned m_jac = __ballot_sync(FULL, req == REQ_JAC);
    const unsigned m_lu = __ballot_sync(FULL, req == REQ_LU);
    if (req == REQ_RHS) {
      const SrcScalars sc{s.src.rhoe_init, s.src.rhoesrc};
      const double* sp_p = s.src.p;
      const long long sp_s = s.src.stride;
      rhs_tab<TabC>(
        s.t_eval, p.reactor_type, s.vec(s.in_sel), s.vec(s.out_sel), s.scratch(), s.delta(), sc,
        [sp_p, sp_s](const int n) { return __ldg(sp_p + n * sp_s); });
      req = REQ_NONE;
    }
    __syncwarp();
    if (m_jac && ((__popc(m_jac) >= sp.jac_defer_min_lanes) || (m_rhs == 0))) {
      if (req == REQ_JAC) {
        double yy[NEQ];
#pragma unroll
        for (int k = 0; k < NEQ; k++) {
          yy[k] = s.y()[k];
        }
        double* J = s.savedJ();
        jac_analytic(p.reactor_type, yy, J, [J](const int row, const int col, const double v) { J[col * NEQ + row] = v; });
        req = REQ_NONE;
      }
      __syncwarp();
    }
    if (m_lu && ((__popc(m_lu) >= sp.lu_defer_min_lanes) || (m_rhs == 0))) {
      if (req == REQ_LU) {
        s.lu_info = (NEQ <= 12) ? dense_getrf_small(s.A(), s.piv(), s.perm()) : ((NEQ <= 24) ? dense_getrf_mid(s.A(), s.piv(), s.perm()) : dense_getrf(s.A(), s.piv()));
        req = REQ_NONE;
      }
      __syncwarp();
    }
    if (req == REQ_DONE) {
      // box_unflatten (ReactorUtils.H:222-291)
      const double rhoe_out = s.src.rhoe_init + a.dt * s.src.rhoesrc;
      const double fc = double(s.nfe + s.nfeDQ);
      if (a.layout == LAYOUT_ARRAY4) {
        const long long o = cell_offset(cell, a.len, a.rY.js, a.rY.ks);
        double rho = 0.0;
        if (p.clean_init_massfrac) {
          double rho_old = 0.0, rho_oldclipped = 0.0;
          for (int k = 0; k < N; k++) {
            rho_old += a.rY.p[o + k * a.rY.ns];
          }
          for (int k = 0; k < N; k++) {
            rho_oldclipped += fmin(fmax(0.0, a.rY.p[o + k * a.rY.ns]), rho_old);
          }
          for (int k = 0; k < N; k++) {
            const double yold = a.rY.p[o + k * a.rY.ns];
            const double clipped = fmin(fmax(0.0, yold), rho_old);
            s.y()[k] = yold + (s.y()[k] - clipped) * rho_old / rho_oldclipped;
          }
        }
#pragma unroll
        for (int k = 0; k < N; k++) {
          const double v = s.y()[k];
          a.rY.p[o + k * a.rY.ns] = v;
          rho += v;
        }
        const double rho_inv = 1.0 / rho;
#pragma unroll
        for (int k = 0; k < N; k++) {
          s.scratch()[k] = s.y()[k] * rho_inv;
        }
        double temp = s.y()[N];
        t_from_energy_tab<TabC>(p.reactor_type, rhoe_out * rho_inv, s.scratch(), temp);
        a.rE.p[cell_offset(cell, a.len, a.rE.js, a.rE.ks)] = rhoe_out;
        a.T.p[cell_offset(cell, a.len, a.T.js, a.T.ks)] = temp;
        a.FC.p[cell_offset(cell, a.len, a.FC.js, a.FC.ks)] = fc;
      } else {
#pragma unroll
        for (int k = 0; k < NEQ; k++) {
          a.y[(long long)cell * NEQ + k] = s.y()[k];
        }
        a.e[cell] = rhoe_out;
        if (a.fc != nullptr) {
          a.fc[cell] = fc;
        }
      }
      ls.cells++;
      ls.steps += s.nst;
      ls.rhs += s.nfe + s.nfeDQ;
      ls.maxrhs = (unsigned long long)(s.nfe + s.nfeDQ) > ls.maxrhs ? (unsigned long long)(s.nfe + s.nfeDQ) : ls.maxrhs;
      ls.jac += s.nje;
      ls.setups += s.nsetups;
      ls.newton += s.nni;
      ls.errfail += s.netf;
      ls.convfail += s.ncfn;
      ls.failed += (s.flag != CV_SUCCESS);
      ls.maxsteps = (unsigned long long)s.nst > ls.maxsteps ? s.nst : ls.maxsteps;
      cell = -1;
      req = REQ_NONE;
    }
  }
  flush_stats(ls, stats);
}

template <bool A_SMEM>
cudaError_t launch_bdf2(const BoxArgs& a, const BdfParams& p, const SchedParams& sp, int* queue, DevStats* stats, cudaStream_t stream)
{
  typedef BdfStoreSmem<32, A_SMEM> Store;
  auto kern = PRX_CAT(bdf2_kernel_, PRX_MECH_SUFFIX)<A_SMEM>;
  const size_t smem = sizeof(double) * 32 * (size_t)Store::doubles_per_thread(p.qmax);
  int dev = 0, sms = 148, per_sm = 0, max_optin = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  if (smem > (size_t)max_optin) {
    return cudaErrorInvalidConfiguration;
  }
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) {
    return e;
  }
  cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 32, smem) != cudaSuccess || per_sm < 1) {
    return cudaErrorInvalidConfiguration;
  }
  if (sp.max_warps_per_sm > 0 && per_sm > sp.max_warps_per_sm) {
    per_sm = sp.max_warps_per_sm;
  }
  long long grid = (long long)sms * per_sm; // persistent: one full wave, a multiple of the SM count
  const long long needed = ((long long)a.ncells + 31) / 32;
  grid = needed < grid ? needed : grid;
  e = cudaMemsetAsync(queue, 0, sizeof(int), stream);
  if (e != cudaSuccess) {
    return e;
  }
  kern<<<(int)grid, 32, smem, stream>>>(a, p, sp, queue, stats);
  return cudaGetLastError();
}
#ifndef PRX_BDF3_MINBLOCKS
#define PRX_BDF3_MINBLOCKS 12
#endif
// hot/cold split (BdfStoreHot); TABRHS: table-driven RHS (rolled) instead of the generated straight-line one
template <bool TABRHS>
__global__ void __launch_bounds__(128, PRX_BDF3_MINBLOCKS / 4) PRX_CAT(bdf3_kernel_, PRX_MECH_SUFFIX)(
  const BoxArgs a, const BdfParams p, const SchedParams sp, int* __restrict__ queue, DevStats* __restrict__ stats,
  double* __restrict__ cold, const int cold_doubles)
{
  typedef BdfCellT<BdfStoreHot<32, TABRHS>, SrcRef> Cell;
  Cell s;
  // blocks of W warps: each warp owns its own 32 columns; W > 1 adds a block barrier per loop iteration that keeps
  // the SM's warps in the same phase (RHS / control), i.e. sharing instruction-cache fills
  const int wib = threadIdx.x >> 5, lane_ = threadIdx.x & 31, nwarp = blockDim.x >> 5;
  s.hot = prx_dyn_smem + (size_t)wib * 32 * Cell::hot_doubles() + lane_;
  s.cold = cold + ((size_t)blockIdx.x * nwarp + wib) * 32 * cold_doubles + lane_;
  s.nz = p.qmax + 1;
  int cell = -1;
  bool exhausted = false;
  int req = REQ_NONE;
  LaneStats ls = {};
  BdfParams pc = p;
  pc.array4_semantics = 0; // the initial T solve is done here with the table-driven Newton loop
  for (;;) {
    // ---- (1) refill ----
    int got = fetch_cell((cell < 0) && !exhausted, queue, a.ncells);
    if (got >= 0 && a.order != nullptr) {
      got = a.order[got]; // cells handed out in sorted order (see prx_api.cu build_order)
    }
    if (got == -2) {
      exhausted = true;
    } else if (got >= 0) {
      double y0[NEQ];
      CellSrc cs;
      if (load_cell(a, got, y0, cs, p.clean_init_massfrac)) {
        cell = got;
        s.src.rhoe_init = cs.rhoe_init;
        s.src.rhoesrc = cs.rhoesrc;
        if (a.layout == LAYOUT_ARRAY4) {
          s.src.p = a.rYsrc.p + cell_offset(got, a.len, a.rYsrc.js, a.rYsrc.ks);
          s.src.stride = a.rYsrc.ns;
        } else {
          s.src.p = a.ysrc + (long long)got * N;
          s.src.stride = 1;
        }
        double rho = 0.0;
#pragma unroll
        for (int k = 0; k < N; k++) {
          s.zn(0)[k] = y0[k];
          rho += y0[k];
        }
        double temp = y0[N];
        if (p.array4_semantics) { // box_flatten re-solves T from (e, Y) (ReactorUtils.H:206-219)
          const double rho_inv = 1.0 / rho;
          if (TABRHS) {
#pragma unroll
            for (int k = 0; k < N; k++) {
              s.scratch()[k] = y0[k] * rho_inv;
            }
            t_from_energy_tab<TabC>(p.reactor_type, cs.rhoe_init * rho_inv, s.scratch(), temp);
          } else {
            double Y[N];
#pragma unroll
            for (int k = 0; k < N; k++) {
              Y[k] = y0[k] * rho_inv;
            }
            t_from_energy(p.reactor_type, cs.rhoe_init * rho_inv, Y, temp);
          }
        }
        s.zn(0)[N] = temp;
        bdf_begin(s, pc, a.dt, a.time);
        req = REQ_NONE;
      } else {
        store_masked(a, got);
        ls.masked++;
      }
    }
    if (nwarp > 1) {
      if (!__syncthreads_or(!(exhausted && (cell < 0)))) {
        break;
      }
    } else if (!__any_sync(FULL, !(exhausted && (cell < 0)))) {
      break;
    }
    // ---- (2) control ----
    if (cell >= 0 && req == REQ_NONE) {
      req = bdf_resume(s, pc);
    }
    __syncwarp();
    // ---- (3) serve ----
    const unsigned m_rhs = __ballot_sync(FULL, req == REQ_RHS);
    const unsigned m_jac = __ballot_sync(FULL, req == REQ_JAC);
    const unsigned m_lu = __ballot_sync(FULL, req == REQ_LU);
    if (req == REQ_RHS) {
      const SrcScalars sc{s.src.rhoe_init, s.src.rhoesrc};
      const double* sp_p = s.src.p;
      const long long sp_s = s.src.stride;
      if (TABRHS) {
        rhs_tab<TabC>(
          s.t_eval, p.reactor_type, s.vec(s.in_sel), s.vec(s.out_sel), s.scratch(), s.delta(), sc,
          [sp_p, sp_s](const int n) { return __ldg(sp_p + n * sp_s); });
      } else {
        double yy[NEQ], yd[NEQ];
        CellSrc cs;
        cs.rhoe_init = sc.rhoe_init;
        cs.rhoesrc = sc.rhoesrc;
        typename Cell::Vec vin = s.vec(s.in_sel);
#pragma unroll
        for (int k = 0; k < NEQ; k++) {
          yy[k] = vin[k];
        }
#pragma unroll
        for (int k = 0; k < N; k++) {
          cs.rYsrc[k] = __ldg(sp_p + k * sp_s);
        }
        rhs(s.t_eval, p.reactor_type, yy, yd, cs);
        typename Cell::Vec vout = s.vec(s.out_sel);
#pragma unroll
        for (int k = 0; k < NEQ; k++) {
          vout[k] = yd[k];
        }
      }
      req = REQ_NONE;
    }
    __syncwarp();
    if (m_jac && ((__popc(m_jac) >= sp.jac_defer_min_lanes) || (m_rhs == 0))) {
      if (req == REQ_JAC) {
        double yy[NEQ];
#pragma unroll
        for (int k = 0; k < NEQ; k++) {
          yy[k] = s.y()[k];
        }
        double* J = s.savedJ();
        jac_analytic(p.reactor_type, yy, J, [J](const int row, const int col, const double v) { J[col * NEQ + row] = v; });
        req = REQ_NONE;
      }
      __syncwarp();
    }
    if (m_lu && ((__popc(m_lu) >= sp.lu_defer_min_lanes) || (m_rhs == 0))) {
      if (req == REQ_LU) {
        s.lu_info = (NEQ <= 12) ? dense_getrf_small(s.A(), s.piv(), s.perm()) : ((NEQ <= 24) ? dense_getrf_mid(s.A(), s.piv(), s.perm()) : dense_getrf(s.A(), s.piv()));
        req = REQ_NONE;
      }
      __syncwarp();
    }
    if (req == REQ_DONE) {
      // box_unflatten (ReactorUtils.H:222-291)
      const double rhoe_out = s.src.rhoe_init + a.dt * s.src.rhoesrc;
      const double fc = double(s.nfe + s.nfeDQ);
      if (a.layout == LAYOUT_ARRAY4) {
        const long long o = cell_offset(cell, a.len, a.rY.js, a.rY.ks);
        double rho = 0.0;
        if (p.clean_init_massfrac) {
          double rho_old = 0.0, rho_oldclipped = 0.0;
          for (int k = 0; k < N; k++) {
            rho_old += a.rY.p[o + k * a.rY.ns];
          }
          for (int k = 0; k < N; k++) {
            rho_oldclipped += fmin(fmax(0.0, a.rY.p[o + k * a.rY.ns]), rho_old);
          }
          for (int k = 0; k < N; k++) {
            const double yold = a.rY.p[o + k * a.rY.ns];
            const double clipped = fmin(fmax(0.0, yold), rho_old);
            s.y()[k] = yold + (s.y()[k] - clipped) * rho_old / rho_oldclipped;
          }
        }
#pragma unroll
        for (int k = 0; k < N; k++) {
          const double v = s.y()[k];
          a.rY.p[o + k * a.rY.ns] = v;
          rho += v;
        }
        const double rho_inv = 1.0 / rho;
        double temp = s.y()[N];
        if (TABRHS) {
#pragma unroll
          for (int k = 0; k < N; k++) {
            s.scratch()[k] = s.y()[k] * rho_inv;
          }
          t_from_energy_tab<TabC>(p.reactor_type, rhoe_out * rho_inv, s.scratch(), temp);
        } else {
          double Y[N];
#pragma unroll
          for (int k = 0; k < N; k++) {
            Y[k] = s.y()[k] * rho_inv;
          }
          t_from_energy(p.reactor_type, rhoe_out * rho_inv, Y, temp);
        }
        a.rE.p[cell_offset(cell, a.len, a.rE.js, a.rE.ks)] = rhoe_out;
        a.T.p[cell_offset(cell, a.len, a.T.js, a.T.ks)] = temp;
        a.FC.p[cell_offset(cell, a.len, a.FC.js, a.FC.ks)] = fc;
      } else {
#pragma unroll
        for (int k = 0; k < NEQ; k++) {
          a.y[(long long)cell * NEQ + k] = s.y()[k];
        }
        a.e[cell] = rhoe_out;
        if (a.fc != nullptr) {
          a.fc[cell] = fc;
        }
      }
      ls.cells++;
      ls.steps += s.nst;
      ls.rhs += s.nfe + s.nfeDQ;
      ls.maxrhs = (unsigned long long)(s.nfe + s.nfeDQ) > ls.maxrhs ? (unsigned long long)(s.nfe + s.nfeDQ) : ls.maxrhs;
      ls.jac += s.nje;
      ls.setups += s.nsetups;
      ls.newton += s.nni;
      ls.errfail += s.netf;
      ls.convfail += s.ncfn;
      ls.failed += (s.flag != CV_SUCCESS);
      ls.maxsteps = (unsigned long long)s.nst > ls.maxsteps ? s.nst : ls.maxsteps;
      cell = -1;
      req = REQ_NONE;
    }
  }
  flush_stats(ls, stats);
}

template <bool TABRHS>
cudaError_t launch_bdf3(const BoxArgs& a, const BdfParams& p, const SchedParams& sp, int* queue, DevStats* stats, cudaStream_t stream)
{
  typedef BdfStoreHot<32, TABRHS> Store;
  auto kern = PRX_CAT(bdf3_kernel_, PRX_MECH_SUFFIX)<TABRHS>;
  const int cold_doubles = Store::cold_doubles(p.qmax);
  int dev = 0, sms = 148, per_sm = 0, max_optin = 0;
  const int W0 = (sp.block_warps >= 1 && sp.block_warps <= 4) ? sp.block_warps : 1;
  const size_t smem = sizeof(double) * 32 * W0 * (size_t)Store::hot_doubles();
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  if (smem > (size_t)max_optin) {
    return cudaErrorInvalidConfiguration;
  }
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) {
    return e;
  }
  cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  const int W = (sp.block_warps >= 1 && sp.block_warps <= 4) ? sp.block_warps : 1;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 32 * W, smem) != cudaSuccess || per_sm < 1) {
    return cudaErrorInvalidConfiguration;
  }
  if (sp.max_warps_per_sm > 0 && per_sm * W > sp.max_warps_per_sm) {
    per_sm = sp.max_warps_per_sm / W > 0 ? sp.max_warps_per_sm / W : 1;
  }
  long long grid = (long long)sms * per_sm; // persistent: one full wave, a multiple of the SM count
  const long long needed = ((long long)a.ncells + 32 * W - 1) / (32 * W);
  grid = needed < grid ? needed : grid;
  double* cold = nullptr; // cold columns of the resident cells (stream-ordered, pool-backed after the first call)
  e = cudaMallocAsync(&cold, sizeof(double) * (size_t)grid * 32 * W * cold_doubles, stream);
  if (e != cudaSuccess) {
    return e;
  }
  e = cudaMemsetAsync(queue, 0, sizeof(int), stream);
  if (e == cudaSuccess) {
    kern<<<(int)grid, 32 * W, smem, stream>>>(a, p, sp, queue, stats, cold, cold_doubles);
    e = cudaGetLastError();
  }
  const cudaError_t e2 = cudaFreeAsync(cold, stream);
  return e != cudaSuccess ? e : e2;
}
#endif // PRX_HAVE_BDF2

// =====================================================================================
// BDF kernel, fourth generation ("fat threads"): the WHOLE integrator state on chip at one cell per thread
// =====================================================================================
// ncu on the first-generation kernel (profiles/r01n_*): 4.7 KB of thread-local state per cell x 21 warps
// per SM = 466 MB of local memory against a 126 MB L2 -> 112 KB of DRAM traffic per cell (417x the algorithmic bytes),
// 41 % of the stall samples on local-memory loads.  On-chip capacity (227 KB of shared memory per SM) holds the state
// of ~130 LiDryer cells, so this kernel runs ONE block of NT = 128 threads per SM with
//  * every integrator vector, the Newton matrix and the external species source in shared-memory columns
//    (element i at base[i*NT + tid]: conflict-free, compile-time offsets),
//  * the straight-line generated RHS / Jacobian in registers (up to 255 per thread: the instruction-level parallelism of
//    the 21 independent reactions replaces the thread-level parallelism the state does not leave room for),
//  * savedJ (written once, read once per linear setup) thread-local,
//  * the same phase barrier / sorted work queue / inline Jacobian + LU service as the first generation.
struct SrcScal
{
  double rhoe_init, rhoesrc;
};

#define PRX_HAVE_BDF4 (PRX_FULL_VARIANTS && PRX_NUM_SPECIES <= 12)
// Variant 6 (round 2, session r02z): the same kernel body with the block shape of the first generation (PRX_BLOCK x
// PRX_MINBLOCKS, straight-line chemistry, Newton matrix and saved Jacobian thread-local) and the VECTORS of the integrator
// (Nordsieck array, ewt, acor, tempv, ftemp, y, delta, species source) as coalesced columns of a GLOBAL buffer that is pinned
// in L2 by an access-policy window (persisting lines, cudaLimitPersistingL2CacheSize): the vectors are what every step and
// every Newton iteration re-reads and re-writes after 20 other warps have flushed them out of an LRU L2; local memory cannot
// carry an access policy, an explicit buffer can.  No occupancy is traded for it.
// MEASURED (r02z, 128^3 LiDryer): 67.5 ms (first generation), 78.1 ms with the vectors in the global buffer, 83.8 ms with the
// buffer pinned in L2 (PRX_L2_PERSIST=1): the persisting carve-out takes the L2 away from the Newton matrix, the saved
// Jacobian and the spills, which lose more than the vectors gain.  A dead end like variants 1-5: built only with
// -DPRX_FULL_VARIANTS=1.
#ifndef PRX_HAVE_BDF6
#define PRX_HAVE_BDF6 (PRX_FULL_VARIANTS && PRX_NUM_SPECIES <= 24 && !defined(PRX_AUDIT_RK64_ONLY))
#endif
#if PRX_HAVE_BDF4 || PRX_HAVE_BDF6
template <int NT, bool A_SMEM, int MINB, bool GLOBAL>
__global__ void __launch_bounds__(NT, MINB) PRX_CAT(bdf4_kernel_, PRX_MECH_SUFFIX)(
  const BoxArgs a, const BdfParams p, const SchedParams sp, int* __restrict__ queue, DevStats* __restrict__ stats,
  double* __restrict__ gbuf, const int doubles_per_thread)
{
  typedef BdfCellT<BdfStoreSmem<NT, A_SMEM>, SrcScal> Cell;
  Cell s;
  s.base = GLOBAL ? (gbuf + (size_t)blockIdx.x * NT * (size_t)doubles_per_thread + threadIdx.x) : (prx_dyn_smem + threadIdx.x);
  s.nz = p.qmax + 1;
  int cell = -1;
  bool exhausted = false;
  int req = REQ_NONE;
  LaneStats ls = {};
  BdfParams pc = p;
  pc.array4_semantics = 0; // the initial T solve is done in the refill below
  for (;;) {
    // ---- (1) refill ----
    int got = fetch_cell((cell < 0) && !exhausted, queue, a.ncells);
    if (got >= 0 && a.order != nullptr) {
      got = a.order[got];
    }
    if (got == -2) {
      exhausted = true;
    } else if (got >= 0) {
      double y0[NEQ];
      CellSrc cs;
      if (load_cell(a, got, y0, cs, p.clean_init_massfrac)) {
        cell = got;
        s.src.rhoe_init = cs.rhoe_init;
        s.src.rhoesrc = cs.rhoesrc;
        double rho = 0.0;
#pragma unroll
        for (int k = 0; k < N; k++) {
          s.zn(0)[k] = y0[k];
          s.scratch()[k] = cs.rYsrc[k];
          rho += y0[k];
        }
        double temp = y0[N];
        if (p.array4_semantics) { // box_flatten re-solves T from (e, Y) (ReactorUtils.H:206-219)
          const double rho_inv = 1.0 / rho;
          double Y[N];
#pragma unroll
          for (int k = 0; k < N; k++) {
            Y[k] = y0[k] * rho_inv;
          }
          t_from_energy(p.reactor_type, cs.rhoe_init * rho_inv, Y, temp);
        }
        s.zn(0)[N] = temp;
        bdf_begin(s, pc, a.dt, a.time);
        req = REQ_NONE;
      } else {
        store_masked(a, got);
        ls.masked++;
      }
    }
    if (!__syncthreads_or(!(exhausted && (cell < 0)))) {
      break;
    }
    // ---- (2) control, Jacobian / LU requests served inside (every iteration ends in one RHS per warp) ----
    for (;;) {
      if (cell >= 0 && req == REQ_NONE) {
        req = bdf_resume(s, pc);
      }
      if (__ballot_sync(FULL, req == REQ_JAC || req == REQ_LU) == 0u) {
        break;
      }
      if (req == REQ_JAC) {
        double yy[NEQ];
#pragma unroll
        for (int k = 0; k < NEQ; k++) {
          yy[k] = s.y()[k];
        }
        double* J = s.savedJ();
        jac_analytic(p.reactor_type, yy, J, [J](const int row, const int col, const double v) { J[col * NEQ + row] = v; });
        req = REQ_NONE;
      }
      if (req == REQ_LU) {
        s.lu_info = (NEQ <= 12) ? dense_getrf_small(s.A(), s.piv(), s.perm()) : ((NEQ <= 24) ? dense_getrf_mid(s.A(), s.piv(), s.perm()) : dense_getrf(s.A(), s.piv()));
        req = REQ_NONE;
      }
    }
    __syncthreads(); // phase barrier: the block's warps walk through the RHS code together (instruction cache)
    // ---- (3) serve ----
    if (req == REQ_RHS) {
      double yy[NEQ], yd[NEQ];
      CellSrc cs;
      cs.rhoe_init = s.src.rhoe_init;
      cs.rhoesrc = s.src.rhoesrc;
      typename Cell::Vec vin = s.vec(s.in_sel);
#pragma unroll
      for (int k = 0; k < NEQ; k++) {
        yy[k] = vin[k];
      }
#pragma unroll
      for (int k = 0; k < N; k++) {
        cs.rYsrc[k] = s.scratch()[k];
      }
      rhs(s.t_eval, p.reactor_type, yy, yd, cs);
      typename Cell::Vec vout = s.vec(s.out_sel);
#pragma unroll
      for (int k = 0; k < NEQ; k++) {
        vout[k] = yd[k];
      }
      req = REQ_NONE;
    }
    if (req == REQ_DONE) {
      double yy[NEQ];
      CellSrc cs;
      cs.rhoe_init = s.src.rhoe_init;
      cs.rhoesrc = s.src.rhoesrc;
#pragma unroll
      for (int k = 0; k < NEQ; k++) {
        yy[k] = s.y()[k];
      }
      store_cell(a, cell, yy, cs, p.reactor_type, double(s.nfe + s.nfeDQ), p.clean_init_massfrac, p.array4_semantics != 0);
      ls.cells++;
      ls.steps += s.nst;
      ls.rhs += s.nfe + s.nfeDQ;
      ls.maxrhs = (unsigned long long)(s.nfe + s.nfeDQ) > ls.maxrhs ? (unsigned long long)(s.nfe + s.nfeDQ) : ls.maxrhs;
      ls.jac += s.nje;
      ls.setups += s.nsetups;
      ls.newton += s.nni;
      ls.errfail += s.netf;
      ls.convfail += s.ncfn;
      ls.failed += (s.flag != CV_SUCCESS);
      ls.maxsteps = (unsigned long long)s.nst > ls.maxsteps ? s.nst : ls.maxsteps;
      cell = -1;
      req = REQ_NONE;
    }
  }
  flush_stats(ls, stats);
}

#endif // PRX_HAVE_BDF4 || PRX_HAVE_BDF6

#if PRX_HAVE_BDF6
cudaError_t launch_bdf6(const BoxArgs& a, const BdfParams& p, const SchedParams& sp, int* queue, DevStats* stats, cudaStream_t stream)
{
  constexpr int NT = PRX_BLOCK;
  typedef BdfStoreSmem<NT, false> Store;
  auto kern = PRX_CAT(bdf4_kernel_, PRX_MECH_SUFFIX)<NT, false, PRX_MINBLOCKS, true>;
  int dev = 0, sms = 148, per_sm = 0, max_persist = 0, max_window = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, dev);
  cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, dev);
  cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxL1);
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, NT, 0) != cudaSuccess || per_sm < 1) {
    return cudaErrorInvalidConfiguration;
  }
  long long grid = (long long)sms * per_sm;
  const long long needed = ((long long)a.ncells + NT - 1) / NT;
  grid = needed < grid ? needed : grid;
  const int dpt = Store::doubles_per_thread(p.qmax);
  const size_t bytes = sizeof(double) * (size_t)grid * NT * (size_t)dpt;
  double* gbuf = nullptr;
  cudaError_t e = cudaMallocAsync((void**)&gbuf, bytes, stream);
  if (e != cudaSuccess) {
    return e;
  }
  // pin the buffer in L2: persisting carve-out as large as the device allows, hit ratio scaled when the buffer is larger
  static const bool want_persist = [] {
    const char* ev = std::getenv("PRX_L2_PERSIST");
    return ev == nullptr || ev[0] != '0';
  }();
  bool window_set = false;
  if (want_persist && max_persist > 0 && max_window > 0) {
    static int carved_dev = -1;
    if (carved_dev != dev) {
      cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, (size_t)max_persist);
      carved_dev = dev;
    }
    cudaStreamAttrValue av;
    std::memset(&av, 0, sizeof(av));
    const size_t win = bytes < (size_t)max_window ? bytes : (size_t)max_window;
    av.accessPolicyWindow.base_ptr = gbuf;
    av.accessPolicyWindow.num_bytes = win;
    const double ratio = (double)max_persist / (double)win;
    av.accessPolicyWindow.hitRatio = ratio >= 1.0 ? 1.0f : (float)ratio;
    av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    av.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    window_set = (cudaStreamSetAttribute(stream, cudaStreamAttributeAccessPolicyWindow, &av) == cudaSuccess);
    if (!window_set) {
      cudaGetLastError();
    }
  }
  e = cudaMemsetAsync(queue, 0, sizeof(int), stream);
  if (e == cudaSuccess) {
    kern<<<(int)grid, NT, 0, stream>>>(a, p, sp, queue, stats, gbuf, dpt);
    e = cudaGetLastError();
  }
  if (window_set) { // the window belongs to this launch only: give the caller's stream back as it was
    cudaStreamAttrValue av;
    std::memset(&av, 0, sizeof(av));
    av.accessPolicyWindow.num_bytes = 0;
    cudaStreamSetAttribute(stream, cudaStreamAttributeAccessPolicyWindow, &av);
  }
  cudaFreeAsync(gbuf, stream);
  return e;
}
#endif

#if PRX_HAVE_BDF4
template <int NT>
cudaError_t launch_bdf4_nt(const BoxArgs& a, const BdfParams& p, const SchedParams& sp, int* queue, DevStats* stats, cudaStream_t stream)
{
  typedef BdfStoreSmem<NT, true> Store;
  auto kern = PRX_CAT(bdf4_kernel_, PRX_MECH_SUFFIX)<NT, true, 1, false>;
  const size_t smem = sizeof(double) * NT * (size_t)Store::doubles_per_thread(p.qmax);
  int dev = 0, sms = 148, per_sm = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) {
    return e;
  }
  cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, NT, smem) != cudaSuccess || per_sm < 1) {
    return cudaErrorInvalidConfiguration;
  }
  long long grid = (long long)sms * per_sm;
  const long long needed = ((long long)a.ncells + NT - 1) / NT;
  grid = needed < grid ? needed : grid;
  e = cudaMemsetAsync(queue, 0, sizeof(int), stream);
  if (e != cudaSuccess) {
    return e;
  }
  kern<<<(int)grid, NT, smem, stream>>>(a, p, sp, queue, stats, nullptr, 0);
  return cudaGetLastError();
}

cudaError_t launch_bdf4(const BoxArgs& a, const BdfParams& p, const SchedParams& sp, int* queue, DevStats* stats, cudaStream_t stream)
{
  // the largest block whose state fits the SM's shared memory (one block per SM)
  int dev = 0, max_optin = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  const size_t per_thread = sizeof(double) * (size_t)BdfStoreSmem<32, true>::doubles_per_thread(p.qmax);
  int nt = (int)((size_t)max_optin / per_thread) / 32 * 32;
  if (sp.block_warps > 0 && sp.block_warps * 32 < nt) {
    nt = sp.block_warps * 32;
  }
  if (nt >= 128) {
    return launch_bdf4_nt<128>(a, p, sp, queue, stats, stream);
  }
  if (nt >= 96) {
    return launch_bdf4_nt<96>(a, p, sp, queue, stats, stream);
  }
  if (nt >= 64) {
    return launch_bdf4_nt<64>(a, p, sp, queue, stats, stream);
  }
  if (nt >= 32) {
    return launch_bdf4_nt<32>(a, p, sp, queue, stats, stream);
  }
  return cudaErrorInvalidConfiguration;
}
#endif // PRX_HAVE_BDF4

// =====================================================================================
// BDF kernel, third generation: G lanes per cell (prx_coop.cuh)
// =====================================================================================
// Same warp loop as above (refill / control / serve) with the GROUP as the unit: the 32/G cells of a warp
// run their integrator coroutines (control flow replicated in the G lanes of a group, uniform inside it) and
// meet at one converged place per request kind.  Blocks are single warps; no block barrier.
#ifdef PRX_AUDIT_RK64_ONLY
#define PRX_HAVE_COOP 0
#else
#define PRX_HAVE_COOP 1
#endif
#if PRX_HAVE_COOP
// -DPRX_COOP_SERIAL_JAC: debugging aid - the analytic Jacobian evaluated by ONE lane with the generated
// straight-line code instead of jac_coop (small mechanisms only)

#ifndef PRX_COOP_MINBLOCKS
#define PRX_COOP_MINBLOCKS 16
#endif

#ifdef PRX_COOP_SERIAL_JAC
__device__ __noinline__ void jac_serial(const int reactor_type, const double* __restrict__ y, double* __restrict__ J)
{
  double yy[NEQ];
  double Jt[NEQ * NEQ];
  for (int k = 0; k < NEQ; k++) {
    yy[k] = y[k];
  }
  jac_analytic(reactor_type, yy, Jt, [J](const int row, const int col, const double v) { J[col * NEQ + row] = v; });
}
#endif

// W warps per block: W > 1 adds two block barriers per loop iteration (exit test, and in front of the RHS) that keep the SM's
// warps in the same phase of the 630 KB kernel - ncu r02c on dodecane_lu: 31 % of the stall samples were "no instruction"
// with independent single-warp blocks.
template <int G, int W>
__global__ void __launch_bounds__(32 * W, (PRX_COOP_MINBLOCKS / W) > 0 ? (PRX_COOP_MINBLOCKS / W) : 1) PRX_CAT(bdfc_kernel_, PRX_MECH_SUFFIX)(
  const BoxArgs a, const BdfParams p, const SchedParams sp, int* __restrict__ queue, DevStats* __restrict__ stats,
  double* __restrict__ jscratch, double* __restrict__ ascratch, const int cell_doubles)
{
  typedef BdfCellT<BdfStoreCoop<G>, SrcRef> Cell;
  constexpr int CPW = 32 / G; // cells per warp
  Cell s;
  const unsigned lane = threadIdx.x & 31u;
  const int grp = (int)lane / G;
  const int wib = (int)(threadIdx.x >> 5);
  const Grp<G>& g = s.g;
  const size_t slot = ((size_t)blockIdx.x * W + wib) * CPW + grp; // resident-cell slot of this group
  s.base = prx_dyn_smem + ((size_t)wib * CPW + grp) * cell_doubles;
  s.gJ = jscratch + slot * (NEQ * NEQ);
  s.gA = (ascratch != nullptr) ? ascratch + slot * (NEQ * NEQ) : nullptr;
  s.nz = p.qmax + 1;
  int cell = -1;
  bool exhausted = false;
  int req = REQ_NONE;
  LaneStats ls = {};
  BdfParams pc = p;
  pc.array4_semantics = 0; // the initial T solve is done here, cooperatively
  for (;;) {
    // ---- (1) refill: the group leaders draw cell indices (one atomic per warp), then broadcast ----
    int got = fetch_cell((cell < 0) && !exhausted && (g.gl == 0), queue, a.ncells);
    if (got >= 0 && a.order != nullptr) {
      got = a.order[got];
    }
    got = __shfl_sync(FULL, got, (int)(lane & ~(unsigned)(G - 1)));
    if (got == -2) {
      exhausted = true;
    } else if (got >= 0) {
      bool live = true;
      if (a.layout == LAYOUT_ARRAY4 && a.mask.p != nullptr) {
        live = a.mask.p[cell_offset(got, a.len, a.mask.js, a.mask.ks)] != -1;
      }
      if (live) {
        cell = got;
        double* z0 = s.zn(0);
        double temp;
        if (a.layout == LAYOUT_ARRAY4) {
          const long long o = cell_offset(got, a.len, a.rY.js, a.rY.ks);
          s.src.p = a.rYsrc.p + cell_offset(got, a.len, a.rYsrc.js, a.rYsrc.ks);
          s.src.stride = a.rYsrc.ns;
          for (int k = g.gl; k < N; k += G) {
            z0[k] = __ldg(a.rY.p + o + k * a.rY.ns);
          }
          temp = a.T.p[cell_offset(got, a.len, a.T.js, a.T.ks)];
          s.src.rhoe_init = a.rE.p[cell_offset(got, a.len, a.rE.js, a.rE.ks)];
          s.src.rhoesrc = a.rEsrc.p[cell_offset(got, a.len, a.rEsrc.js, a.rEsrc.ks)];
        } else {
          s.src.p = a.ysrc + (long long)got * N;
          s.src.stride = 1;
          for (int k = g.gl; k < N; k += G) {
            z0[k] = a.y[(long long)got * NEQ + k];
          }
          temp = a.y[(long long)got * NEQ + N];
          s.src.rhoe_init = a.e[got];
          s.src.rhoesrc = a.esrc[got];
        }
        if (a.layout == LAYOUT_ARRAY4) { // box_flatten: optional clipping, then T from (e, Y)
          double part = 0.0;
          for (int k = g.gl; k < N; k += G) {
            part += z0[k];
          }
          double rho = g.sum(part);
          if (p.clean_init_massfrac) { // ReactorUtils.H:183-199: clip to [0, rho_old], density of the clipped state
            part = 0.0;
            for (int k = g.gl; k < N; k += G) {
              z0[k] = fmin(fmax(0.0, z0[k]), rho);
              part += z0[k];
            }
            rho = g.sum(part);
          }
          const double rho_inv = 1.0 / rho;
          for (int k = g.gl; k < N; k += G) {
            s.sc()[k] = z0[k] * rho_inv;
          }
          double cv_unused;
          if (p.reactor_type == 1) {
            t_from_energy_coop<TabG, 1, G>(g, s.src.rhoe_init * rho_inv, s.sc(), temp, cv_unused);
          } else {
            t_from_energy_coop<TabG, 2, G>(g, s.src.rhoe_init * rho_inv, s.sc(), temp, cv_unused);
          }
        }
        if (s.owns(N)) {
          z0[N] = temp;
        }
        g.sync();
        bdf_begin(s, pc, a.dt, a.time);
        req = REQ_NONE;
      } else {
        if (g.gl == 0) {
          store_masked(a, got);
          ls.masked++;
        }
      }
    }
    if (W > 1) {
      if (!__syncthreads_or(!(exhausted && (cell < 0)))) {
        break;
      }
    } else if (!__any_sync(FULL, !(exhausted && (cell < 0)))) {
      break;
    }
    // ---- (2) control ----
    if (cell >= 0 && req == REQ_NONE) {
      req = bdf_resume(s, pc);
    }
    if (W > 1) { // (making this barrier optional at run time - PRX_COOP_BARRIERS, session r02n - cost 23 % by itself and the
                 // one-barrier mode was no faster: unconditional)
      __syncthreads(); // phase barrier: the block's warps enter the RHS / Jacobian / LU service together
    } else {
      __syncwarp();
    }
    // ---- (3) serve ----
    const unsigned m_rhs = __ballot_sync(FULL, req == REQ_RHS);
    const unsigned m_jac = __ballot_sync(FULL, req == REQ_JAC);
    const unsigned m_lu = __ballot_sync(FULL, req == REQ_LU);
    if (req == REQ_RHS) {
      const SrcScalars sc{s.src.rhoe_init, s.src.rhoesrc};
      const double* sp_p = s.src.p;
      const long long sp_s = s.src.stride;
      rhs_coop<TabG, G>(
        g, s.t_eval, p.reactor_type, s.vec(s.in_sel), s.vec(s.out_sel), s.sc(), s.gs(), s.qv(), sc,
        [sp_p, sp_s](const int n) { return __ldg(sp_p + n * sp_s); });
      req = REQ_NONE;
    }
    __syncwarp();
    if (m_jac && ((__popc(m_jac) >= sp.jac_defer_min_lanes * G) || (m_rhs == 0))) {
      if (req == REQ_JAC) {
        g.sync();
#ifdef PRX_COOP_SERIAL_JAC
        if (g.gl == 0) {
          jac_serial(p.reactor_type, s.y(), s.savedJ());
        }
        g.sync();
#else
        jac_coop<TabG, G>(g, p.reactor_type, s.y(), s.A(), s.sc(), s.gs(), s.dq(), s.qv(), s.savedJ());
#endif
        req = REQ_NONE;
      }
      __syncwarp();
    }
    if (m_lu && ((__popc(m_lu) >= sp.lu_defer_min_lanes * G) || (m_rhs == 0))) {
      if (req == REQ_LU) {
#ifndef PRX_COOP_GENERIC_LA
        if (G == 32 && NEQ > 32 && NEQ <= 64) {
          s.lu_info = getrf_warp(s.A(), s.piv(), s.perm()); // one warp per matrix, rows in registers (prx_coop.cuh)
        } else
#endif
        {
          s.lu_info = getrf_coop<G>(g, s.A(), s.piv());
        }
        req = REQ_NONE;
      }
      __syncwarp();
    }
    if (req == REQ_DONE) {
      // box_unflatten (ReactorUtils.H:222-291)
      g.sync();
      const double rhoe_out = s.src.rhoe_init + a.dt * s.src.rhoesrc;
      const double fc = double(s.nfe + s.nfeDQ);
      double* yv = s.y();
      if (a.layout == LAYOUT_ARRAY4) {
        const long long o = cell_offset(cell, a.len, a.rY.js, a.rY.ks);
        if (p.clean_init_massfrac) { // increment measured against the clipped input (ReactorUtils.H:244-271)
          double p0 = 0.0;
          for (int k = g.gl; k < N; k += G) {
            p0 += a.rY.p[o + k * a.rY.ns];
          }
          const double rho_old = g.sum(p0);
          double p1 = 0.0;
          for (int k = g.gl; k < N; k += G) {
            p1 += fmin(fmax(0.0, a.rY.p[o + k * a.rY.ns]), rho_old);
          }
          const double rho_oldclipped = g.sum(p1);
          for (int k = g.gl; k < N; k += G) {
            const double yold = a.rY.p[o + k * a.rY.ns];
            const double clipped = fmin(fmax(0.0, yold), rho_old);
            yv[k] = yold + (yv[k] - clipped) * rho_old / rho_oldclipped;
          }
        }
        double part = 0.0;
        for (int k = g.gl; k < N; k += G) {
          const double v = yv[k];
          a.rY.p[o + k * a.rY.ns] = v;
          part += v;
        }
        const double rho = g.sum(part);
        const double rho_inv = 1.0 / rho;
        for (int k = g.gl; k < N; k += G) {
          s.sc()[k] = yv[k] * rho_inv;
        }
        double temp = yv[N];
        double cv_unused;
        if (p.reactor_type == 1) {
          t_from_energy_coop<TabG, 1, G>(g, rhoe_out * rho_inv, s.sc(), temp, cv_unused);
        } else {
          t_from_energy_coop<TabG, 2, G>(g, rhoe_out * rho_inv, s.sc(), temp, cv_unused);
        }
        if (g.gl == 0) {
          a.rE.p[cell_offset(cell, a.len, a.rE.js, a.rE.ks)] = rhoe_out;
          a.T.p[cell_offset(cell, a.len, a.T.js, a.T.ks)] = temp;
          a.FC.p[cell_offset(cell, a.len, a.FC.js, a.FC.ks)] = fc;
        }
      } else {
        for (int k = g.gl; k < NEQ; k += G) {
          a.y[(long long)cell * NEQ + k] = yv[k];
        }
        if (g.gl == 0) {
          a.e[cell] = rhoe_out;
          if (a.fc != nullptr) {
            a.fc[cell] = fc;
          }
        }
      }
      if (g.gl == 0) {
        ls.cells++;
        ls.steps += s.nst;
        ls.rhs += s.nfe + s.nfeDQ;
        ls.maxrhs = (unsigned long long)(s.nfe + s.nfeDQ) > ls.maxrhs ? (unsigned long long)(s.nfe + s.nfeDQ) : ls.maxrhs;
        ls.jac += s.nje;
        ls.setups += s.nsetups;
        ls.newton += s.nni;
        ls.errfail += s.netf;
        ls.convfail += s.ncfn;
        ls.failed += (s.flag != CV_SUCCESS);
        ls.maxsteps = (unsigned long long)s.nst > ls.maxsteps ? s.nst : ls.maxsteps;
      }
      cell = -1;
      req = REQ_NONE;
    }
  }
  flush_stats(ls, stats);
}

template <int G, int W>
cudaError_t launch_bdfc_w(const BoxArgs& a, const BdfParams& p, const SchedParams& sp, int* queue, DevStats* stats, cudaStream_t stream)
{
  auto kern = PRX_CAT(bdfc_kernel_, PRX_MECH_SUFFIX)<G, W>;
  constexpr int CPW = 32 / G;
  // Newton matrix in the shared-memory block unless it is large (PRX_COOP_A_GLOBAL: 0 never, 1 always, else > 8 KB)
  const bool a_global = (sp.coop_a_global == 1) || (sp.coop_a_global != 0 && sizeof(double) * NEQ * NEQ > 8192);
  const int cell_doubles = BdfStoreCoop<G>::doubles_per_cell(p.qmax, a_global);
  const size_t smem = sizeof(double) * (size_t)W * CPW * cell_doubles;
  int dev = 0, sms = 148, per_sm = 0, max_optin = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  if (smem > (size_t)max_optin) {
    return cudaErrorInvalidConfiguration;
  }
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) {
    return e;
  }
  cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 32 * W, smem) != cudaSuccess || per_sm < 1) {
    return cudaErrorInvalidConfiguration;
  }
  if (sp.max_warps_per_sm > 0 && per_sm * W > sp.max_warps_per_sm) {
    per_sm = sp.max_warps_per_sm / W > 0 ? sp.max_warps_per_sm / W : 1;
  }
  long long grid = (long long)sms * per_sm; // persistent: one full wave, a multiple of the SM count
  const long long needed = ((long long)a.ncells + CPW * W - 1) / (CPW * W);
  grid = needed < grid ? needed : grid;
  double* jscratch = nullptr; // savedJ slots, one per resident cell (stream-ordered, pool-backed after the first call)
  const size_t slots = (size_t)grid * W * CPW;
  e = cudaMallocAsync(&jscratch, sizeof(double) * slots * NEQ * NEQ * (a_global ? 2 : 1), stream);
  if (e != cudaSuccess) {
    return e;
  }
  double* ascratch = a_global ? jscratch + slots * NEQ * NEQ : nullptr;
  e = cudaMemsetAsync(queue, 0, sizeof(int), stream);
  if (e == cudaSuccess) {
    kern<<<(int)grid, 32 * W, smem, stream>>>(a, p, sp, queue, stats, jscratch, ascratch, cell_doubles);
    e = cudaGetLastError();
  }
  const cudaError_t e2 = cudaFreeAsync(jscratch, stream);
  return e != cudaSuccess ? e : e2;
}

#ifndef PRX_COOP_BLOCK_WARPS
#define PRX_COOP_BLOCK_WARPS 1 // default warps per block of the cooperative kernel (PRX_BLOCK_WARPS overrides: 1 or 4)
#endif
template <int G>
cudaError_t launch_bdfc(const BoxArgs& a, const BdfParams& p, const SchedParams& sp, int* queue, DevStats* stats, cudaStream_t stream)
{
  const int w = sp.block_warps > 0 ? sp.block_warps : PRX_COOP_BLOCK_WARPS;
  if (w >= 8 && G == 32) { // (W > 1 is built for the one-warp-per-cell layout only: compile time)
    return launch_bdfc_w<(G == 32 ? G : 32), 8>(a, p, sp, queue, stats, stream);
  }
  if (w >= 4 && G == 32) {
    return launch_bdfc_w<(G == 32 ? G : 32), 4>(a, p, sp, queue, stats, stream);
  }
  return launch_bdfc_w<G, 1>(a, p, sp, queue, stats, stream);
}
#endif // PRX_HAVE_COOP

// =====================================================================================
// probes (parity tests of the device thermo-chemistry)
// =====================================================================================
__global__ void PRX_CAT(probe_rhs_kernel_, PRX_MECH_SUFFIX)(
  const int n, const double t, const int reactor_type, const double* __restrict__ y,
  const double* __restrict__ rhoe, const double* __restrict__ rhoesrc, const double* __restrict__ rYsrc,
  double* __restrict__ ydot)
{
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) {
    return;
  }
  double yy[NEQ], yd[NEQ];
  CellSrc src;
  for (int k = 0; k < NEQ; k++) {
    yy[k] = y[(long long)s * NEQ + k];
  }
  for (int k = 0; k < N; k++) {
    src.rYsrc[k] = rYsrc[(long long)s * N + k];
  }
  src.rhoe_init = rhoe[s];
  src.rhoesrc = rhoesrc[s];
  rhs(t, reactor_type, yy, yd, src);
  for (int k = 0; k < NEQ; k++) {
    ydot[(long long)s * NEQ + k] = yd[k];
  }
}

__global__ void PRX_CAT(probe_jac_kernel_, PRX_MECH_SUFFIX)(
  const int n, const int reactor_type, const double* __restrict__ y, double* __restrict__ J,
  double* __restrict__ scratch)
{
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) {
    return;
  }
  double yy[NEQ];
  for (int k = 0; k < NEQ; k++) {
    yy[k] = y[(long long)s * NEQ + k];
  }
  double* Jt = scratch + (long long)s * NEQ * NEQ;
  double* Jo = J + (long long)s * NEQ * NEQ;
#if PRX_HAVE_BDF1
  jac_analytic(reactor_type, yy, Jt, [Jo](const int row, const int col, const double v) { Jo[col * NEQ + row] = v; });
#else
  jac_analytic_tab<TabG>(reactor_type, yy, Jt, [Jo](const int row, const int col, const double v) { Jo[col * NEQ + row] = v; });
#endif
}

// =====================================================================================
// ReactorBase::flatten / unflatten (FlattenOps, PP/Source/Reactions/ReactorUtils.H:160-355):
// Array4 box <-> flat vectors, ordering 0 = YCOrder (cell-major), 1 = CYOrder (component-major).
// One thread per cell; the Array4 side is coalesced along i.
// =====================================================================================
__device__ __forceinline__ long long vec_index(const int ordering, const int comp, const int icell, const int ncells, const int nc)
{
  return ordering == 0 ? (long long)icell * nc + comp : (long long)comp * ncells + icell;
}

__global__ void PRX_CAT(flatten_kernel_, PRX_MECH_SUFFIX)(
  const BoxArgs a, const int reactor_type, const int clean_init_massfrac, const int ordering, double* __restrict__ y_vect,
  double* __restrict__ src_vect, double* __restrict__ vect_energy, double* __restrict__ src_vect_energy)
{
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= a.ncells) {
    return;
  }
  double y[NEQ];
  CellSrc src;
  BoxArgs b = a;
  b.mask.p = nullptr; // flatten knows no mask
  load_cell(b, c, y, src, clean_init_massfrac);
  double rho = 0.0, Y[N];
  for (int k = 0; k < N; k++) {
    rho += y[k];
  }
  const double rho_inv = 1.0 / rho;
  for (int k = 0; k < N; k++) {
    Y[k] = y[k] * rho_inv;
    y_vect[vec_index(ordering, k, c, a.ncells, NEQ)] = y[k];
    src_vect[vec_index(ordering, k, c, a.ncells, N)] = src.rYsrc[k];
  }
  vect_energy[c] = src.rhoe_init;
  src_vect_energy[c] = src.rhoesrc;
  double temp = y[N];
  t_from_energy(reactor_type, src.rhoe_init * rho_inv, Y, temp);
  y_vect[vec_index(ordering, N, c, a.ncells, NEQ)] = temp;
}

__global__ void PRX_CAT(unflatten_kernel_, PRX_MECH_SUFFIX)(
  const BoxArgs a, const int reactor_type, const int clean_init_massfrac, const int ordering,
  const double* __restrict__ y_vect, const double* __restrict__ vect_energy, const long long* __restrict__ fcunt)
{
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= a.ncells) {
    return;
  }
  double y[NEQ];
  for (int k = 0; k < NEQ; k++) {
    y[k] = y_vect[vec_index(ordering, k, c, a.ncells, NEQ)];
  }
  CellSrc src;
  src.rhoe_init = vect_energy[c];
  src.rhoesrc = a.rEsrc.p[cell_offset(c, a.len, a.rEsrc.js, a.rEsrc.ks)];
  store_cell(a, c, y, src, reactor_type, (double)fcunt[c], clean_init_massfrac, true);
}

// ---- unfused react_state pack / unpack (for integrator kernels whose load/store phase is not fused: cooperative kernel) ----
__global__ void PRX_CAT(state_pack_kernel_, PRX_MECH_SUFFIX)(const BoxArgs a)
{
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= a.ncells) {
    return;
  }
  const double* so = a.sold.p + cell_offset(c, a.len, a.sold.js, a.sold.ks);
  const long long o = cell_offset(c, a.len, a.rY.js, a.rY.ks);
  for (int k = 0; k < N; k++) {
    a.rY.p[o + k * a.rY.ns] = so[(a.ix.ufs + k) * a.sold.ns];
  }
  a.T.p[cell_offset(c, a.len, a.T.js, a.T.ks)] = so[a.ix.utemp * a.sold.ns];
  a.rE.p[cell_offset(c, a.len, a.rE.js, a.rE.ks)] = so[a.ix.ueint * a.sold.ns];
  a.rEsrc.p[cell_offset(c, a.len, a.rEsrc.js, a.rEsrc.ks)] = state_rhoedot_ext(a, c);
}

__global__ void PRX_CAT(state_unpack_kernel_, PRX_MECH_SUFFIX)(const BoxArgs a)
{
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= a.ncells) {
    return;
  }
  double rhoY[N];
  const long long o = cell_offset(c, a.len, a.rY.js, a.rY.ks);
  for (int k = 0; k < N; k++) {
    rhoY[k] = a.rY.p[o + k * a.rY.ns];
  }
  state_unpack(
    a, c, rhoY, a.T.p[cell_offset(c, a.len, a.T.js, a.T.ks)], a.rEsrc.p[cell_offset(c, a.len, a.rEsrc.js, a.rEsrc.ks)],
    a.FC.p[cell_offset(c, a.len, a.FC.js, a.FC.ks)]);
}

// ---- PeleC::computeTemp's pointwise kernel pc_cmpTemp (PC/Source/Utilities.H:70-84): T <- REY2T(rho, e, Y) in
// place on a state FAB.  One thread per cell, every component read is a coalesced 256-byte warp transaction along i; the
// kernel is HBM-bound ((N + 3) doubles in, 1 out per cell).  (reset_internal_energy, the dual-energy bookkeeping PeleC runs
// first, stays with the caller.)
__global__ void PRX_CAT(compute_temp_kernel_, PRX_MECH_SUFFIX)(const BoxArgs a)
{
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < a.ncells; c += gridDim.x * blockDim.x) {
    double* S = a.snew.p + cell_offset(c, a.len, a.snew.js, a.snew.ks);
    const long long ns = a.snew.ns;
    const double rho = S[a.ix.urho * ns];
    const double rhoInv = 1.0 / rho;
    double T = S[a.ix.utemp * ns];
    const double e = S[a.ix.ueint * ns] * rhoInv;
    double Y[N];
#pragma unroll
    for (int n = 0; n < N; n++) {
      Y[n] = __ldg(S + (a.ix.ufs + n) * ns) * rhoInv;
    }
    t_from_energy<1>(e, Y, T);
    S[a.ix.utemp * ns] = T;
  }
}

// Same, two cells (consecutive in i) per thread with 128-bit loads / stores: every component row is read as double2
// (north_star: "SoA state is read with coalesced, vectorised HBM loads").  Used when the rows are 16-byte aligned and of even
// length (launch_compute_temp checks); small mechanisms only (two cells of mass fractions in registers).
#define PRX_HAVE_CT2 (PRX_NUM_SPECIES <= 24)
#if PRX_HAVE_CT2
__global__ void PRX_CAT(compute_temp2_kernel_, PRX_MECH_SUFFIX)(const BoxArgs a)
{
  const int npair = a.ncells >> 1;
  for (int pr = blockIdx.x * blockDim.x + threadIdx.x; pr < npair; pr += gridDim.x * blockDim.x) {
    double* S = a.snew.p + cell_offset(pr << 1, a.len, a.snew.js, a.snew.ks); // len[0] even: both cells on one row
    const long long ns = a.snew.ns;
    const double2 rho = *reinterpret_cast<const double2*>(S + a.ix.urho * ns);
    const double2 re = *reinterpret_cast<const double2*>(S + a.ix.ueint * ns);
    double2 T = *reinterpret_cast<const double2*>(S + a.ix.utemp * ns);
    const double ia = 1.0 / rho.x, ib = 1.0 / rho.y;
    double Ya[N], Yb[N];
#pragma unroll
    for (int n = 0; n < N; n++) {
      const double2 v = __ldg(reinterpret_cast<const double2*>(S + (a.ix.ufs + n) * ns));
      Ya[n] = v.x * ia;
      Yb[n] = v.y * ib;
    }
    t_from_energy<1>(re.x * ia, Ya, T.x);
    t_from_energy<1>(re.y * ib, Yb, T.y);
    *reinterpret_cast<double2*>(S + a.ix.utemp * ns) = T;
  }
}
#endif

cudaError_t launch_compute_temp(const BoxArgs& a, cudaStream_t stream)
{
  if (a.ncells <= 0) {
    return cudaSuccess;
  }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const long long cap = (long long)sms * 8; // a multiple of the SM count; grid-stride beyond it
#if PRX_HAVE_CT2
  // Measured (session r02o, 256^3 FAB): 0.843 vs 0.823 ms (LiDryer), 2.09 vs 1.73 ms (drm19) - the kernel is bound by issue
  // slots / the FP64 pipe (T-Newton), not by its loads, and two cells per thread cost registers: the 128-bit form is OFF
  // unless PRX_CT_VEC2=1 (tests/test_gpu_state.py exercises both).
  static const bool want_vec2 = [] {
    const char* e = std::getenv("PRX_CT_VEC2");
    return e != nullptr && e[0] == '1';
  }();
  const bool vec2 = want_vec2 && (a.len[0] % 2 == 0) && (reinterpret_cast<unsigned long long>(a.snew.p) % 16 == 0) && (a.snew.js % 2 == 0) &&
                    (a.snew.ks % 2 == 0) && (a.snew.ns % 2 == 0);
  if (vec2) {
    const long long want = ((long long)(a.ncells >> 1) + 255) / 256;
    PRX_CAT(compute_temp2_kernel_, PRX_MECH_SUFFIX)<<<(int)(want < cap ? want : cap), 256, 0, stream>>>(a);
    return cudaGetLastError();
  }
#endif
  const long long want = ((long long)a.ncells + 255) / 256;
  PRX_CAT(compute_temp_kernel_, PRX_MECH_SUFFIX)<<<(int)(want < cap ? want : cap), 256, 0, stream>>>(a);
  return cudaGetLastError();
}

// ---- PeleC's conserved -> primitive conversion pc_ctoprim (PC/Source/Utilities.H:242-321) --------------
// One thread per cell, grid-stride; every component row is a coalesced warp transaction along i.  Per cell: NVAR doubles
// in, QVAR + NQAUX out (Fuego EOS: ideal-gas mixture, T from (e, Y) by the reference's Newton iteration).  Each EOS call is
// restated in the reference's own association (RTY2Cs sums Y_k cv_k with CKCVMS, RTY2G / RTY2dpde_dpdre use CKCVBS).
__global__ void __launch_bounds__(256, (PRX_NUM_SPECIES <= 16) ? 4 : 2) PRX_CAT(ctoprim_kernel_, PRX_MECH_SUFFIX)(const PrimArgs a)
{
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < a.ncells; c += gridDim.x * blockDim.x) {
    const double* U = a.u.p + cell_offset(c, a.len, a.u.js, a.u.ks);
    double* Q = a.q.p + cell_offset(c, a.len, a.q.js, a.q.ks);
    double* QA = a.qa.p + cell_offset(c, a.len, a.qa.js, a.qa.ks);
    const long long un = a.u.ns, qn = a.q.ns, an = a.qa.ns;
    const double rho = __ldg(U + a.ix.urho * un);
    const double rhoinv = 1.0 / rho;
    const double vx = __ldg(U + a.ix.umx * un) * rhoinv;
    const double vy = __ldg(U + a.ix.umy * un) * rhoinv;
    const double vz = __ldg(U + a.ix.umz * un) * rhoinv;
    const double kineng = 0.5 * rho * (vx * vx + vy * vy + vz * vz) - 0.5 * rho * 0.0 * 0.0 * 0.0 * 0.0; // omega = rad = 0
    Q[a.qx.qrho * qn] = rho;
    Q[a.qx.qu * qn] = vx;
    Q[a.qx.qv * qn] = vy;
    Q[a.qx.qw * qn] = vz;
    double Y[N];
    const double tiny = 1e-4 * 2.220446049250313e-16;
#pragma unroll
    for (int n = 0; n < N; n++) {
      double y = __ldg(U + (a.ix.ufs + n) * un) / rho;
      if (-tiny < y && y < tiny) {
        y = 0.0;
      }
      Y[n] = y;
      Q[(a.qx.qfs + n) * qn] = y;
    }
    const double e = (__ldg(U + a.ix.ueden * un) - kineng) * rhoinv;
    double T = __ldg(U + a.ix.utemp * un);
    double yow = 0.0; // Y2WBAR / CKMMWY / CKPY all accumulate Y_k * imw_k in species order
#pragma unroll
    for (int n = 0; n < N; n++) {
      yow += Y[n] * prx_imw[n];
    }
    const double wbar = 1.0 / yow;
    t_from_energy<1>(e, Y, T);
    const double p = rho * PRX_RU * T * yow;          // CKPY
    double cvms[N];
    cv_R(cvms, T);                                    // CKCVMS = cv_R * (RU / mw_k literal)
    double cv_s = 0.0;
#pragma unroll
    for (int n = 0; n < N; n++) {
      cv_s += Y[n] * (cvms[n] * tr_ru_mw[n]);
    }
    const double g_s = (wbar * cv_s + PRX_RU) / (wbar * cv_s); // RTY2Cs
    const double cs = sqrt(g_s * p / rho);
    const double cvx = wbar * ckcvbs(T, Y);           // TY2G
    const double gam1 = (cvx + PRX_RU) / cvx;
    Q[a.qx.qtemp * qn] = T;
    Q[a.qx.qreint * qn] = e * rho;
    Q[a.qx.qpres * qn] = p;
    Q[a.qx.qgame * qn] = p / (e * rho) + 1.0;
    QA[a.qx.qdpdr * an] = p / rho;
    QA[a.qx.qdpde * an] = (gam1 - 1.0) * rho;
    QA[a.qx.qgamc * an] = gam1;
    QA[a.qx.qc * an] = cs;
    QA[a.qx.qcsml * an] = fmax(a.small_num, a.small_num * cs);
    QA[a.qx.qrspec * an] = PRX_RU / wbar;
  }
}

cudaError_t launch_ctoprim(const PrimArgs& a, cudaStream_t stream)
{
  if (a.ncells <= 0) {
    return cudaSuccess;
  }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const long long cap = (long long)sms * 8, want = ((long long)a.ncells + 255) / 256;
  PRX_CAT(ctoprim_kernel_, PRX_MECH_SUFFIX)<<<(int)(want < cap ? want : cap), 256, 0, stream>>>(a);
  return cudaGetLastError();
}

// ---- SimpleTransport::get_transport_coeffs (PP/Source/Transport/Simple.H:230-447 with the Fuego EOS; PeleC's call site
// PC/Source/Diffusion.cpp:182-222): mixture-averaged rho*D_k, shear viscosity mu, bulk viscosity xi, conductivity lambda from
// (T, rho, Y) with the mechanism's log-T fits.  One thread per cell.  The binary-diffusion sum needs exp(-fit_ij(T)) for every
// ordered species pair; the fit table is symmetric bit for bit (checked by the generator), so each unordered pair is
// evaluated once and added to both species in the reference's own summation order (ascending partner index).
#ifdef PRX_HAVE_TRANSPORT
#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif
__global__ void __launch_bounds__(128) PRX_CAT(transport_kernel_, PRX_MECH_SUFFIX)(const PrimArgs a)
{
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < a.ncells; c += gridDim.x * blockDim.x) {
    const double* Yp = a.Y.p + cell_offset(c, a.len, a.Y.js, a.Y.ks);
    const double Tloc = a.T.p[cell_offset(c, a.len, a.T.js, a.T.ks)];
    double Y[N], X[N];
    double sum = 0.0;
#pragma unroll
    for (int i = 0; i < N; i++) {
      Y[i] = __ldg(Yp + i * a.Y.ns);
      sum += Y[i];
    }
    const double logT0 = log(Tloc), logT1 = logT0 * logT0, logT2 = logT0 * logT1;
#pragma unroll
    for (int i = 0; i < N; i++) {
      Y[i] += 1.e-15 * (sum / N - Y[i]);
    }
    double wbar = 0.0;
#pragma unroll
    for (int i = 0; i < N; i++) {
      wbar += Y[i] * tr_iwt[i];
    }
    wbar = 1.0 / wbar;
#pragma unroll
    for (int i = 0; i < N; i++) {
      X[i] = Y[i] * wbar * tr_iwt[i];
    }
    // shear + bulk viscosity
    double mu = 0.0, xi = 0.0;
    {
      double cvk[N];
      cv_R(cvk, Tloc); // CKCVMS = cv_R * (RU / mw_k literal)
      const double pi3_2 = pow(M_PI, 1.5);
#pragma unroll 1
      for (int i = 0; i < N; i++) {
        double m = tr_fitmu[4 * i] + tr_fitmu[1 + 4 * i] * logT0 + tr_fitmu[2 + 4 * i] * logT1 + tr_fitmu[3 + 4 * i] * logT2;
        m = exp(m);
        mu += X[i] * (m * m * m * m * m * m);
        const int nlin = tr_nlin[i];
        if (nlin != 0) {
          const double cvms = cvk[i] * tr_ru_mw[i];
          const double cvkint = cvms * tr_wt[i] / PRX_RU - 1.50;
          const double cvkrot = (nlin == 1) ? 1.0 : 1.5;
          const double e298 = tr_eps[i] / 298.0, eT = tr_eps[i] / Tloc;
          const double Fnorm = 1.0 + 0.50 * pi3_2 * sqrt(e298) + (2.0 + .50 * M_PI * M_PI) * e298 + pi3_2 * sqrt(e298) * e298;
          const double FofT = 1.0 + 0.50 * pi3_2 * sqrt(eT) + (2.0 + .50 * M_PI * M_PI) * eT + pi3_2 * sqrt(eT) * eT;
          const double xl = 0.250 * M_PI * (cvkint / (cvkint + 1.50) * cvkint / (cvkint + 1.50)) * tr_zrot[i] / cvkrot * Fnorm / FofT * m;
          xi += X[i] * (sqrt(sqrt(xl * xl * xl)));
        }
      }
      mu = cbrt(sqrt(mu));
      xi = cbrt(xi * xi * xi * xi);
    }
    // conductivity
    double lam = 0.0;
#pragma unroll 1
    for (int i = 0; i < N; i++) {
      double l = tr_fitlam[4 * i] + tr_fitlam[1 + 4 * i] * logT0 + tr_fitlam[2 + 4 * i] * logT1 + tr_fitlam[3 + 4 * i] * logT2;
      l = exp(l);
      lam += X[i] * sqrt(sqrt(l));
    }
    lam = lam * lam * lam * lam;
    // mixture-averaged diffusion coefficients
    double term2[N];
#pragma unroll
    for (int i = 0; i < N; i++) {
      term2[i] = 0.0;
    }
// rolled pair loops (term2[] in L1-resident local memory): unrolling them for small mechanisms was measured slower
// (r02y, 192^3 LiDryer: 2.48 vs 2.37 ms)
#define PRX_TR_PAIR_UNROLL _Pragma("unroll 1")
#if PRX_TR_DBIN_SYMMETRIC
    PRX_TR_PAIR_UNROLL
    for (int i = 0; i < N; i++) {
      PRX_TR_PAIR_UNROLL
      for (int j = i + 1; j < N; j++) {
        const int f = 4 * (i + N * j);
        const double d = tr_fitdbin[f] + tr_fitdbin[1 + f] * logT0 + tr_fitdbin[2 + f] * logT1 + tr_fitdbin[3 + f] * logT2;
        const double ex = exp(-d);
        term2[i] += X[j] * ex;
        term2[j] += X[i] * ex;
      }
    }
#else
#pragma unroll 1
    for (int i = 0; i < N; i++) {
#pragma unroll 1
      for (int j = 0; j < N; j++) {
        if (i != j) {
          const int f = 4 * (i + N * j);
          const double d = tr_fitdbin[f] + tr_fitdbin[1 + f] * logT0 + tr_fitdbin[2 + f] * logT1 + tr_fitdbin[3 + f] * logT2;
          term2[i] += X[j] * exp(-d);
        }
      }
    }
#endif
    const double scale = 1.01325e+06 / (PRX_RU * Tloc); // Constants::PATM / (RU T)
    double* Dp = a.D.p + cell_offset(c, a.len, a.D.js, a.D.ks);
#pragma unroll
    for (int i = 0; i < N; i++) {
      double term1 = 0.0;
#pragma unroll
      for (int j = 0; j < N; j++) {
        if (j != i) {
          term1 += Y[j];
        }
      }
      Dp[i * a.D.ns] = tr_wt[i] * term1 / term2[i] * scale;
    }
    a.mu.p[cell_offset(c, a.len, a.mu.js, a.mu.ks)] = mu;
    a.xi.p[cell_offset(c, a.len, a.xi.js, a.xi.ks)] = xi;
    a.lam.p[cell_offset(c, a.len, a.lam.js, a.lam.ks)] = lam;
  }
}

cudaError_t launch_transport(const PrimArgs& a, cudaStream_t stream)
{
  if (a.ncells <= 0) {
    return cudaSuccess;
  }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const long long cap = (long long)sms * 16, want = ((long long)a.ncells + 127) / 128;
  PRX_CAT(transport_kernel_, PRX_MECH_SUFFIX)<<<(int)(want < cap ? want : cap), 128, 0, stream>>>(a);
  return cudaGetLastError();
}
#else
cudaError_t launch_transport(const PrimArgs&, cudaStream_t) { return cudaErrorNotSupported; }
#endif

// ---- ReactorArkode stage (PP/Source/Reactions/ReactorArkode.cpp:385-420 cF_RHS, SUN/src/arkode/arkode_erkstep.c:650-700):
// ycur = sum_j coef_j F_j + yn (the order N_VLinearCombination adds them), then F_is = f(t, ycur) for every cell.  One thread
// per cell, CYOrder columns: every read and write is a coalesced row.
__global__ void __launch_bounds__(PRX_RK_BLOCK, PRX_RK_MINBLOCKS) PRX_CAT(ark_stage_kernel_, PRX_MECH_SUFFIX)(const ArkStageArgs a)
{
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= a.ncells) {
    return;
  }
  const long long nc = a.ncells;
  double y[NEQ], yd[NEQ];
#pragma unroll
  for (int k = 0; k < NEQ; k++) {
    const long long o = k * nc + c;
    if (a.nterms > 0) {
      double acc = a.coef[0] * a.F[0][o];
      for (int j = 1; j < a.nterms; j++) {
        acc += a.coef[j] * a.F[j][o];
      }
      acc += 1.0 * a.yn[o];
      y[k] = acc;
      if (a.ycur != nullptr) { // the stage state is only needed here, in registers
        a.ycur[o] = acc;
      }
    } else {
      y[k] = a.yn[o];
    }
  }
  CellSrc src;
  src.rhoe_init = a.rhoe[c];
  src.rhoesrc = a.rhoesrc[c];
#pragma unroll
  for (int k = 0; k < N; k++) {
    src.rYsrc[k] = a.rYsrc[k * nc + c];
  }
  rhs(a.t, a.reactor_type, y, yd, src);
#pragma unroll
  for (int k = 0; k < NEQ; k++) {
    a.fout[k * nc + c] = yd[k];
  }
}

cudaError_t launch_ark_stage(const ArkStageArgs& a, cudaStream_t stream)
{
  if (a.ncells <= 0) {
    return cudaSuccess;
  }
  PRX_CAT(ark_stage_kernel_, PRX_MECH_SUFFIX)<<<(a.ncells + PRX_RK_BLOCK - 1) / PRX_RK_BLOCK, PRX_RK_BLOCK, 0, stream>>>(a);
  return cudaGetLastError();
}

cudaError_t launch_state_pack(const BoxArgs& a, cudaStream_t stream)
{
  if (a.ncells <= 0) {
    return cudaSuccess;
  }
  PRX_CAT(state_pack_kernel_, PRX_MECH_SUFFIX)<<<(a.ncells + 127) / 128, 128, 0, stream>>>(a);
  return cudaGetLastError();
}

cudaError_t launch_state_unpack(const BoxArgs& a, const prx_config&, cudaStream_t stream)
{
  if (a.ncells <= 0) {
    return cudaSuccess;
  }
  PRX_CAT(state_unpack_kernel_, PRX_MECH_SUFFIX)<<<(a.ncells + 127) / 128, 128, 0, stream>>>(a);
  return cudaGetLastError();
}

cudaError_t launch_flatten(
  const BoxArgs& a, const prx_config& cfg, int ordering, double* y, double* ysrc, double* e, double* esrc, cudaStream_t stream)
{
  if (a.ncells <= 0) {
    return cudaSuccess;
  }
  PRX_CAT(flatten_kernel_, PRX_MECH_SUFFIX)<<<(a.ncells + 127) / 128, 128, 0, stream>>>(
    a, cfg.reactor_type, cfg.clean_init_massfrac, ordering, y, ysrc, e, esrc);
  return cudaGetLastError();
}

cudaError_t launch_unflatten(
  const BoxArgs& a, const prx_config& cfg, int ordering, const double* y, const double* e, const long long* fcunt, cudaStream_t stream)
{
  if (a.ncells <= 0) {
    return cudaSuccess;
  }
  PRX_CAT(unflatten_kernel_, PRX_MECH_SUFFIX)<<<(a.ncells + 127) / 128, 128, 0, stream>>>(
    a, cfg.reactor_type, cfg.clean_init_massfrac, ordering, y, e, fcunt);
  return cudaGetLastError();
}

// =====================================================================================
// launchers
// =====================================================================================
int persistent_grid(const void* kernel, const int ncells, const int block = PRX_BLOCK, const int dyn_smem = 0)
{
  int dev = 0, sms = 148, per_sm = 1;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, block, dyn_smem) != cudaSuccess || per_sm < 1) {
    per_sm = 1;
  }
  long long grid = (long long)sms * per_sm; // a multiple of the SM count: one full wave, persistent
  const long long needed = ((long long)ncells + block - 1) / block;
  if (needed < grid) {
    grid = needed;
  }
  return (int)(grid < 1 ? 1 : grid);
}

cudaError_t launch_rk64(const BoxArgs& a, const prx_config& cfg, int* queue, DevStats* stats, cudaStream_t stream)
{
  if (a.ncells <= 0) {
    return cudaSuccess;
  }
  Rk64Params p;
  p.reactor_type = cfg.reactor_type;
  p.atol = cfg.atol;
  p.nsub_guess = cfg.rk64_nsubsteps_guess;
  p.nsub_min = cfg.rk64_nsubsteps_min;
  p.nsub_max = cfg.rk64_nsubsteps_max;
  p.array4_semantics = (a.layout != LAYOUT_FLAT);
  auto kern = PRX_CAT(rk64_kernel_, PRX_MECH_SUFFIX);
  const int grid = persistent_grid((const void*)kern, a.ncells, PRX_RK_BLOCK);
  cudaError_t e = cudaMemsetAsync(queue, 0, sizeof(int), stream);
  if (e != cudaSuccess) {
    return e;
  }
  // ReactorRK64 reacts every cell of the box: the reference never reads the mask (ReactorRK64.cpp:175).  load_cell would
  // skip a masked cell without loading it, and the integrator would run on the lane's previous cell.
  BoxArgs b = a;
  b.mask.p = nullptr;
  kern<<<grid, PRX_RK_BLOCK, 0, stream>>>(b, p, queue, stats);
  return cudaGetLastError();
}

int effective_bdf_variant(const SchedParams& sp);

cudaError_t launch_bdf(
  const BoxArgs& a, const prx_config& cfg, const double* typ_vals, int* queue, DevStats* stats,
  const SchedParams& sp, cudaStream_t stream)
{
  if (a.ncells <= 0) {
    return cudaSuccess;
  }
  BdfParams p;
  p.reactor_type = cfg.reactor_type;
  p.rtol = cfg.rtol;
  for (int k = 0; k < NEQ; k++) {
    p.atolv[k] = (typ_vals != nullptr && typ_vals[0] > 0.0) ? typ_vals[k] * cfg.atol : cfg.atol;
  }
  p.qmax = cfg.max_order;
  p.maxcor = cfg.max_nls_iters;
  p.analytic_jac = (cfg.solve_type == PRX_SOLVE_DENSE_AJ_DIRECT);
  p.mxstep = 10000;
  p.maxnef = 100;
  p.msbj = 100;
  p.clean_init_massfrac = cfg.clean_init_massfrac;
  p.array4_semantics = (a.layout != LAYOUT_FLAT);
  // kernel generation: 1 = shared-memory columns incl. the Newton matrix, 2 = columns, matrix thread-local,
  // 0 = first generation (thread-local state, straight-line chemistry); default per mechanism
  const int variant = effective_bdf_variant(sp);
#if PRX_HAVE_COOP
  switch (variant) { // G lanes per cell
  case 8:
    return launch_bdfc<8>(a, p, sp, queue, stats, stream);
  case 16:
    return launch_bdfc<16>(a, p, sp, queue, stats, stream);
  case 32:
    return launch_bdfc<32>(a, p, sp, queue, stats, stream);
  default:
    break;
  }
#endif
#if PRX_HAVE_BDF2
  if (variant == 1) {
    return launch_bdf2<true>(a, p, sp, queue, stats, stream);
  }
  if (variant == 2) {
    return launch_bdf2<false>(a, p, sp, queue, stats, stream);
  }
  if (variant == 3) {
    return launch_bdf3<false>(a, p, sp, queue, stats, stream);
  }
  if (variant == 4) {
    return launch_bdf3<true>(a, p, sp, queue, stats, stream);
  }
#endif
#if PRX_HAVE_BDF6
  if (variant == 6 && a.layout != LAYOUT_STATE) {
    return launch_bdf6(a, p, sp, queue, stats, stream);
  }
#endif
#if PRX_HAVE_BDF4
  if (variant == 5) {
    return launch_bdf4(a, p, sp, queue, stats, stream);
  }
#endif
#if PRX_HAVE_BDF1
  auto kern = PRX_CAT(bdf_kernel_, PRX_MECH_SUFFIX);
#if PRX_SMEM_SRC
  // the shared-memory columns of PRX_MINBLOCKS resident blocks (+ 1 KB per block reserved by the system) must fit the
  // carve-out, the rest stays L1 for the thread-local state
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, PRX_BDF1_SMEM_BYTES);
  cudaFuncSetAttribute(
    kern, cudaFuncAttributePreferredSharedMemoryCarveout,
    (int)((100LL * PRX_MINBLOCKS * (PRX_BDF1_SMEM_BYTES + 1024) + 228 * 1024 - 1) / (228 * 1024)) + 2);
#else
  cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxL1); // state lives in L1-cached local memory
#endif
  int grid = persistent_grid((const void*)kern, a.ncells, PRX_BLOCK, PRX_BDF1_SMEM_BYTES);
  if (sp.max_warps_per_sm > 0) { // optional cap on resident warps (fewer resident cells: smaller L2 footprint)
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int blocks = sp.max_warps_per_sm / (PRX_BLOCK / 32) > 0 ? sp.max_warps_per_sm / (PRX_BLOCK / 32) : 1;
    grid = grid < sms * blocks ? grid : sms * blocks;
  }
  cudaError_t e = cudaMemsetAsync(queue, 0, sizeof(int), stream);
  if (e != cudaSuccess) {
    return e;
  }
  kern<<<grid, PRX_BLOCK, PRX_BDF1_SMEM_BYTES, stream>>>(a, p, sp, queue, stats);
  return cudaGetLastError();
#elif PRX_HAVE_COOP
  return launch_bdfc<32>(a, p, sp, queue, stats, stream);
#else
  (void)queue;
  (void)stats;
  (void)stream;
  (void)p;
  return cudaErrorNotSupported; // audit build: no BDF kernels
#endif
}

#define PRX_STR2(x) #x
#define PRX_STR(x) PRX_STR2(x)
int effective_bdf_variant(const SchedParams& sp)
{
  int v = sp.bdf_variant < 0 ? PRX_BDF_DEFAULT_VARIANT : sp.bdf_variant;
#if !PRX_HAVE_COOP
  if (v >= 8) {
    v = 0;
  }
#endif
#if !PRX_HAVE_BDF2
  if (v >= 1 && v <= 4) {
    v = 0;
  }
#endif
#if !PRX_HAVE_BDF4
  if (v == 5) {
    v = 0;
  }
#else
  (void)0;
#endif
#if !PRX_HAVE_BDF6
  if (v == 6) {
    v = 0;
  }
#endif
#if !PRX_HAVE_BDF1
  if (v == 0) {
    v = 32;
  }
#endif
  return v;
}

const char* kernel_name(int reactor_kind, const SchedParams& sp)
{
  if (reactor_kind == PRX_REACTOR_RK64) {
    return "rk64_kernel_" PRX_STR(PRX_MECH_SUFFIX);
  }
  if (reactor_kind == PRX_REACTOR_ARKODE) {
    return "ark_stage_kernel_" PRX_STR(PRX_MECH_SUFFIX);
  }
  switch (effective_bdf_variant(sp)) {
  case 8:
  case 16:
  case 32:
    return "bdfc_kernel_" PRX_STR(PRX_MECH_SUFFIX);
  case 1:
  case 2:
    return "bdf2_kernel_" PRX_STR(PRX_MECH_SUFFIX);
  case 3:
  case 4:
    return "bdf3_kernel_" PRX_STR(PRX_MECH_SUFFIX);
  case 5:
  case 6:
    return "bdf4_kernel_" PRX_STR(PRX_MECH_SUFFIX);
  default:
    return "bdf_kernel_" PRX_STR(PRX_MECH_SUFFIX);
  }
}

cudaError_t probe_rhs(
  int n, double t, int reactor_type, const double* y, const double* rhoe, const double* rhoesrc,
  const double* rYsrc, double* ydot, cudaStream_t stream)
{
  PRX_CAT(probe_rhs_kernel_, PRX_MECH_SUFFIX)<<<(n + 63) / 64, 64, 0, stream>>>(
    n, t, reactor_type, y, rhoe, rhoesrc, rYsrc, ydot);
  return cudaGetLastError();
}

cudaError_t probe_jac(int n, int reactor_type, const double* y, double* J, cudaStream_t stream)
{
  double* scratch = nullptr;
  cudaError_t e = cudaMalloc(&scratch, sizeof(double) * (size_t)n * NEQ * NEQ);
  if (e != cudaSuccess) {
    return e;
  }
  PRX_CAT(probe_jac_kernel_, PRX_MECH_SUFFIX)<<<(n + 63) / 64, 64, 0, stream>>>(n, reactor_type, y, J, scratch);
  e = cudaGetLastError();
  cudaStreamSynchronize(stream);
  cudaFree(scratch);
  return e;
}

int bdf_is_fused(const SchedParams& sp) { return effective_bdf_variant(sp) == 0; }

const MechOps g_ops = {PRX_NUM_SPECIES, PRX_NUM_REACTIONS, PRX_MECH_NAME, prx_species_names, launch_rk64, launch_bdf, kernel_name, probe_rhs, probe_jac,
                       launch_flatten, launch_unflatten, launch_state_pack, launch_state_unpack, bdf_is_fused, launch_compute_temp,
                       launch_ctoprim, launch_transport, launch_ark_stage};

} // namespace

const MechOps* PRX_CAT(mech_ops_, PRX_MECH_SUFFIX)() { return &g_ops; }

} // namespace prx
