// prx_coop.cuh - COOPERATIVE (G lanes per cell) building blocks of the BDF reactor.
//
// Why: one cell per thread cannot keep a CVODE-sized state (11 vectors + the Newton matrix, 1.7 KB for
// LiDryer ... 28 KB for dodecane_lu) on chip for more than 4-7 warps per SM, and ncu shows that design
// latency-bound (profiles/r01b_ncu_bdf2_lidryer_64.csv: 6 % warps active, issue slot 12 % busy, FP64 pipe
// 4 %).  Here G = 8 / 16 / 32 lanes share one cell:
//   * per-cell state is a contiguous block of shared memory (vectors, species scratch, reaction rates, the
//     Newton matrix), 32/G cells per warp, so 24-32 warps per SM stay resident;
//   * vector statements of the integrator are ONE SIMD statement (lane i owns elements i, i+G, ...),
//     norms are xor-butterfly shuffles (bit-identical in every lane, so the control flow - replicated in the
//     G lanes - stays uniform inside a group);
//   * the RHS evaluates reactions lane-parallel (lane r, r+G, ...) from global-memory mechanism tables
//     (per-lane index, L1/L2 resident), species sums by shuffle reduction, and the net production rates by
//     a per-species GATHER over the reaction-rate vector in the reference's accumulation order;
//   * the dense LU is staged in shared memory: pivot search by warp shuffle arg-max (ties to the lowest row,
//     = SUNDlsMat_denseGETRF's choice), row swap / column scaling / rank-1 update spread over the lanes.
// Arithmetic per element is the reference's; only the ORDER of the species/vector sums differs (tree instead
// of left-to-right), which is within the FMA-contraction noise the parity bars already account for.
//
// Reference functions restated: see prx_tchem.cuh (thermo-chemistry), prx_integrators.cuh (CVODE),
// SUN/src/sundials/sundials_dense.c:82-190 (GETRF/GETRS).
#ifndef PRX_COOP_CUH
#define PRX_COOP_CUH

#include "prx_integrators.cuh"
#include "prx_tchem.cuh"

namespace prx {

template <int G>
struct Grp
{
  int gl;      // lane within the group
  unsigned gm; // lanes of the group
  __device__ __forceinline__ Grp()
  {
    const unsigned lane = threadIdx.x & 31u;
    gl = (int)(lane & (unsigned)(G - 1));
    gm = (G == 32) ? 0xffffffffu : (((1u << (G & 31)) - 1u) << (lane & ~(unsigned)(G - 1)));
  }
  __device__ __forceinline__ double sum(double v) const
  {
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) {
      v += __shfl_xor_sync(gm, v, o);
    }
    return v;
  }
  __device__ __forceinline__ double max(double v) const
  {
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) {
      const double w = __shfl_xor_sync(gm, v, o);
      v = (w > v) ? w : v;
    }
    return v;
  }
  __device__ __forceinline__ int any(const int b) const { return (__ballot_sync(gm, b != 0) & gm) != 0u; }
  __device__ __forceinline__ void sync() const { __syncwarp(gm); }
};

// ---- storage policy of the integrator state machine (prx_integrators.cuh) -------------------------
// One contiguous block of shared memory per cell:
//   [ (nz + 6) vectors of NEQ | sc N | gs N | dq N | qv NR | piv NEQ ints | perm NEQ ints | A NEQ*NEQ ]
// savedJ (read once per linear setup) lives in a global scratch slot of the group (L2 resident).
template <int G>
struct BdfStoreCoop
{
  typedef double* Vec;
  typedef double* Mat;
  Grp<G> g;
  double* base;
  double* gJ; // savedJ slot (global)
  double* gA; // Newton matrix slot in global memory, or nullptr: the matrix is in the shared-memory block
  int nz;
  // a_global: large systems (dodecane_lu: 23 KB per matrix) keep the Newton matrix in a global (L2 resident)
  // slot instead, which quadruples the number of resident warps; its accesses are coalesced column segments
  __host__ __device__ static int doubles_per_cell(const int qmax, const bool a_global)
  {
    int d = (qmax + 1 + 6) * NEQ + 3 * N + PRX_NUM_REACTIONS + 2 * ((NEQ + 1) / 2) + (a_global ? 0 : NEQ * NEQ);
    if (G < 32) { // cells of one warp start 8 doubles apart modulo 16: their G-wide rows use distinct banks
      while ((d & 15) != 8) {
        d++;
      }
    }
    return d;
  }
  PRX_MEM Vec slot(const int s) const { return base + s * NEQ; }
  PRX_MEM Vec zn(const int j) const { return slot(j); }
  PRX_MEM Vec ewt() const { return slot(nz); }
  PRX_MEM Vec acor() const { return slot(nz + 1); }
  PRX_MEM Vec tempv() const { return slot(nz + 2); }
  PRX_MEM Vec ftemp() const { return slot(nz + 3); }
  PRX_MEM Vec y() const { return slot(nz + 4); }
  PRX_MEM Vec delta() const { return slot(nz + 5); }
  PRX_MEM double* sc() const { return base + (nz + 6) * NEQ; }
  PRX_MEM double* gs() const { return sc() + N; }
  PRX_MEM double* dq() const { return gs() + N; }
  PRX_MEM double* qv() const { return dq() + N; }
  PRX_MEM int* piv() const { return reinterpret_cast<int*>(qv() + PRX_NUM_REACTIONS); }
  PRX_MEM int* perm() const { return piv() + 2 * ((NEQ + 1) / 2); } // composed row permutation (one-warp-per-matrix solve)
  PRX_MEM Mat A() const { return gA != nullptr ? gA : (qv() + PRX_NUM_REACTIONS + 2 * ((NEQ + 1) / 2)); }
  PRX_MEM double* savedJ() const { return gJ; }
  // cooperative hooks
  PRX_MEM int vbegin() const { return g.gl; }
  PRX_MEM int vstep() const { return G; }
  PRX_MEM bool owns(const int i) const { return (i & (G - 1)) == g.gl; }
  PRX_MEM double gsum(const double x) const { return g.sum(x); }
  PRX_MEM double gmax(const double x) const { return g.max(x); }
  PRX_MEM int gor(const int x) const { return g.any(x); }
  PRX_MEM void sync() const { g.sync(); }
  PRX_MEM void getrs(const double* a, const int* p, double* b) const;
};

// ---- dense LU / solve, G lanes per matrix ---------------------------------------------------------
// Same pivot choice (first row of maximal magnitude) and the same per-entry operations as
// SUNDlsMat_denseGETRF / denseGETRS; `a` is column-major NEQ x NEQ, in the cell's shared-memory block or - for
// large systems - in a global (L2 resident) slot.  Lane gl owns rows gl, gl+G, ...  Memory accesses are issued in
// batches of PRX_COOP_CB columns (all loads of a batch before its stores): with the matrix in global memory the
// first version paid one dependent L2 round trip per entry of the trailing update and per column of a triangular
// solve (ncu r01h: 54 % of all stall samples on those lines).
#ifndef PRX_COOP_CB
#define PRX_COOP_CB 4
#endif

template <int G>
__device__ __forceinline__ int getrf_coop(const Grp<G>& g, double* __restrict__ a, int* __restrict__ piv)
{
  constexpr int RPL = (NEQ + G - 1) / G; // rows per lane
  constexpr int CB = PRX_COOP_CB;
  g.sync();
#pragma unroll 1
  for (int k = 0; k < NEQ; k++) {
    double* ck = a + k * NEQ;
    double colv[RPL];
#pragma unroll
    for (int r = 0; r < RPL; r++) {
      const int i = g.gl + r * G;
      colv[r] = (i >= k && i < NEQ) ? ck[i] : 0.0;
    }
    double best = -1.0;
    int bi = k;
#pragma unroll
    for (int r = 0; r < RPL; r++) {
      const int i = g.gl + r * G;
      const double v = fabs(colv[r]);
      if (i >= k && i < NEQ && v > best) {
        best = v;
        bi = i;
      }
    }
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) {
      const double ob = __shfl_xor_sync(g.gm, best, o);
      const int oi = __shfl_xor_sync(g.gm, bi, o);
      if (ob > best || (ob == best && oi < bi)) {
        best = ob;
        bi = oi;
      }
    }
    const int l = (best < 0.0) ? k : bi; // all-NaN column: the reference keeps l = k
    if (g.gl == 0) {
      piv[k] = l;
    }
    if (best == 0.0) {
      g.sync();
      return k + 1;
    }
    if (l != k) {
#pragma unroll 1
      for (int j = g.gl; j < NEQ; j += G) {
        const double t = a[j * NEQ + l];
        a[j * NEQ + l] = a[j * NEQ + k];
        a[j * NEQ + k] = t;
      }
    }
    g.sync();
    const double mult = 1.0 / ck[k];
    double m[RPL]; // this lane's multipliers (column k below the diagonal), kept in registers for the update
#pragma unroll
    for (int r = 0; r < RPL; r++) {
      const int i = g.gl + r * G;
      m[r] = 0.0;
      if (i > k && i < NEQ) {
        m[r] = ck[i] * mult;
        ck[i] = m[r];
      }
    }
    // rank-1 update of the trailing block, CB columns per batch
#pragma unroll 1
    for (int j0 = k + 1; j0 < NEQ; j0 += CB) {
      double akj[CB], x[CB][RPL];
#pragma unroll
      for (int c = 0; c < CB; c++) {
        const int j = (j0 + c < NEQ) ? j0 + c : NEQ - 1;
        akj[c] = a[j * NEQ + k];
#pragma unroll
        for (int r = 0; r < RPL; r++) {
          const int i = g.gl + r * G;
          x[c][r] = (i > k && i < NEQ) ? a[j * NEQ + i] : 0.0;
        }
      }
#pragma unroll
      for (int c = 0; c < CB; c++) {
        if (j0 + c < NEQ && akj[c] != 0.0) {
#pragma unroll
          for (int r = 0; r < RPL; r++) {
            const int i = g.gl + r * G;
            if (i > k && i < NEQ) {
              a[(j0 + c) * NEQ + i] = x[c][r] - akj[c] * m[r];
            }
          }
        }
      }
    }
    g.sync();
  }
  return 0;
}

template <int G>
__device__ __forceinline__ void getrs_coop(
  const Grp<G>& g, const double* __restrict__ a, const int* __restrict__ piv, double* __restrict__ b)
{
  constexpr int RPL = (NEQ + G - 1) / G;
  constexpr int CB = PRX_COOP_CB;
  g.sync();
  if (g.gl == 0) { // row interchanges, in order
#pragma unroll 1
    for (int k = 0; k < NEQ; k++) {
      const int pk = piv[k];
      if (pk != k) {
        const double t = b[k];
        b[k] = b[pk];
        b[pk] = t;
      }
    }
  }
  g.sync();
  // L y = b: this lane's rows of CB columns are fetched together, then the CB eliminations run on shared memory
#pragma unroll 1
  for (int k0 = 0; k0 < NEQ - 1; k0 += CB) {
    double lv[CB][RPL];
#pragma unroll
    for (int c = 0; c < CB; c++) {
      const int k = k0 + c;
#pragma unroll
      for (int r = 0; r < RPL; r++) {
        const int i = g.gl + r * G;
        lv[c][r] = (k < NEQ - 1 && i > k && i < NEQ) ? a[k * NEQ + i] : 0.0;
      }
    }
#pragma unroll
    for (int c = 0; c < CB; c++) {
      const int k = k0 + c;
      if (k < NEQ - 1) {
        const double bk = b[k];
#pragma unroll
        for (int r = 0; r < RPL; r++) {
          const int i = g.gl + r * G;
          if (i > k && i < NEQ) {
            b[i] -= lv[c][r] * bk;
          }
        }
        g.sync();
      }
    }
  }
  // U x = y, from the last column; every lane forms b[k]/a[k][k] itself, the owner of row k stores it
#pragma unroll 1
  for (int k0 = NEQ - 1; k0 >= 0; k0 -= CB) {
    double uv[CB][RPL], dv[CB];
#pragma unroll
    for (int c = 0; c < CB; c++) {
      const int k = k0 - c;
      dv[c] = (k >= 0) ? a[k * NEQ + k] : 1.0;
#pragma unroll
      for (int r = 0; r < RPL; r++) {
        const int i = g.gl + r * G;
        uv[c][r] = (k >= 0 && i < k) ? a[k * NEQ + i] : 0.0;
      }
    }
#pragma unroll
    for (int c = 0; c < CB; c++) {
      const int k = k0 - c;
      if (k >= 0) {
        const double bk = b[k] / dv[c];
#pragma unroll
        for (int r = 0; r < RPL; r++) {
          const int i = g.gl + r * G;
          if (i < k) {
            b[i] -= uv[c][r] * bk;
          }
        }
        g.sync();
        if ((k & (G - 1)) == g.gl) {
          b[k] = bk;
        }
      }
    }
  }
  g.sync();
}

// ---- one WARP per matrix (G = 32): lean forms ------------------------------------------------------------------
// ncu r02h on dodecane_lu (54 x 54): getrf_coop + getrs_coop were 66 % of the executed instructions - 75 k warp instructions
// per factorisation for 3.3 k useful FMAs per lane, 6.7 k per solve for 180 - most of it predicates, clamps and 64-bit index
// arithmetic of the generic batched loops, plus a shared-memory round trip and a group barrier per elimination step.  Here
// lane l owns rows l and l + 32 in REGISTERS for the duration of a column / a solve; the pivot row and the solution entries
// travel by shuffle; columns are walked with running pointers; the row interchanges of a solve are one gather through the
// composed permutation the factorisation leaves behind.  Same pivot choice (first row of maximal magnitude), same per-entry
// operations in the same order as getrf_coop / getrs_coop (= SUNDlsMat_denseGETRF / denseGETRS, sundials_dense.c:82-190).
__device__ __forceinline__ int getrf_warp(double* __restrict__ a, int* __restrict__ piv, int* __restrict__ perm)
{
  static_assert(NEQ <= 64, "two rows per lane");
  constexpr unsigned FULLW = 0xffffffffu;
  const int l = (int)(threadIdx.x & 31u);
  const int r1 = l + 32;
  const bool has1 = r1 < NEQ;
  __syncwarp();
  for (int i = l; i < NEQ; i += 32) {
    perm[i] = i;
  }
  __syncwarp();
#pragma unroll 1
  for (int k = 0; k < NEQ; k++) {
    double* ck = a + k * NEQ;
    const double c0 = (l >= k && l < NEQ) ? ck[l] : 0.0;
    const double c1 = (has1 && r1 >= k) ? ck[r1] : 0.0;
    // arg-max of |column k| over rows >= k, ties to the lowest row
    double best = -1.0;
    int bi = k;
    if (l >= k && l < NEQ) {
      best = fabs(c0);
      bi = l;
    }
    if (has1 && r1 >= k && fabs(c1) > best) {
      best = fabs(c1);
      bi = r1;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double ob = __shfl_xor_sync(FULLW, best, o);
      const int oi = __shfl_xor_sync(FULLW, bi, o);
      if (ob > best || (ob == best && oi < bi)) {
        best = ob;
        bi = oi;
      }
    }
    const int p = (best < 0.0) ? k : bi; // all-NaN column: the reference keeps l = k
    if (l == 0) {
      piv[k] = p;
    }
    if (best == 0.0) {
      __syncwarp();
      return k + 1;
    }
    if (p != k) { // interchange rows k and p in every column (lane j: columns j, j + 32) and in the composed permutation
      for (int j = l; j < NEQ; j += 32) {
        const double t = a[j * NEQ + p];
        a[j * NEQ + p] = a[j * NEQ + k];
        a[j * NEQ + k] = t;
      }
      if (l == 0) {
        const int t = perm[k];
        perm[k] = perm[p];
        perm[p] = t;
      }
      __syncwarp();
    }
    // multipliers of this lane's rows (column k below the diagonal, after the interchange)
    const double akk = ck[k];
    const double mult = 1.0 / akk;
    double v0 = c0, v1 = c1; // this lane's entries of column k after the interchange
    if (p != k) {
      v0 = (l >= k && l < NEQ) ? ck[l] : 0.0;
      v1 = (has1 && r1 >= k) ? ck[r1] : 0.0;
    }
    const bool u0 = (l > k) && (l < NEQ), u1 = has1 && (r1 > k);
    const double m0 = u0 ? v0 * mult : 0.0, m1 = u1 ? v1 * mult : 0.0;
    if (u0) {
      ck[l] = m0;
    }
    if (u1) {
      ck[r1] = m1;
    }
    // rank-1 update of the trailing columns, PRX_LU_CB at a time (loads of a batch before its stores)
#ifndef PRX_LU_CB
#define PRX_LU_CB 4
#endif
    double* cj = ck + NEQ;
    int j = k + 1;
#pragma unroll 1
    for (; j + PRX_LU_CB <= NEQ; j += PRX_LU_CB, cj += PRX_LU_CB * NEQ) {
      double akj[PRX_LU_CB], x0[PRX_LU_CB], x1[PRX_LU_CB];
#pragma unroll
      for (int c = 0; c < PRX_LU_CB; c++) {
        akj[c] = cj[c * NEQ + k];
        x0[c] = u0 ? cj[c * NEQ + l] : 0.0;
        x1[c] = u1 ? cj[c * NEQ + r1] : 0.0;
      }
#pragma unroll
      for (int c = 0; c < PRX_LU_CB; c++) {
        if (akj[c] != 0.0) {
          if (u0) {
            cj[c * NEQ + l] = x0[c] - akj[c] * m0;
          }
          if (u1) {
            cj[c * NEQ + r1] = x1[c] - akj[c] * m1;
          }
        }
      }
    }
#pragma unroll 1
    for (; j < NEQ; j++, cj += NEQ) {
      const double akj = cj[k];
      if (akj != 0.0) {
        if (u0) {
          cj[l] = cj[l] - akj * m0;
        }
        if (u1) {
          cj[r1] = cj[r1] - akj * m1;
        }
      }
    }
    __syncwarp();
  }
  return 0;
}

__device__ __forceinline__ void getrs_warp(const double* __restrict__ a, const int* __restrict__ perm, double* __restrict__ b)
{
  constexpr unsigned FULLW = 0xffffffffu;
  constexpr int CB = 6;
  const int l = (int)(threadIdx.x & 31u);
  const int r1 = l + 32;
  const bool has0 = l < NEQ, has1 = r1 < NEQ;
  __syncwarp();
  // all row interchanges at once: b <- P b through the composed permutation
  double b0 = has0 ? b[perm[l]] : 0.0;
  double b1 = has1 ? b[perm[r1]] : 0.0;
  // L y = b
#pragma unroll 1
  for (int k0 = 0; k0 < NEQ - 1; k0 += CB) {
    double l0[CB], l1[CB];
#pragma unroll
    for (int c = 0; c < CB; c++) {
      const int k = k0 + c;
      const bool in = k < NEQ - 1;
      l0[c] = (in && has0 && l > k) ? a[k * NEQ + l] : 0.0;
      l1[c] = (in && has1 && r1 > k) ? a[k * NEQ + r1] : 0.0;
    }
#pragma unroll
    for (int c = 0; c < CB; c++) {
      const int k = k0 + c;
      if (k < NEQ - 1) {
        const double bk = __shfl_sync(FULLW, (k < 32) ? b0 : b1, k & 31);
        if (has0 && l > k) {
          b0 -= l0[c] * bk;
        }
        if (has1 && r1 > k) {
          b1 -= l1[c] * bk;
        }
      }
    }
  }
  // U x = y, from the last column
#pragma unroll 1
  for (int k0 = NEQ - 1; k0 >= 0; k0 -= CB) {
    double u0[CB], u1[CB], dv[CB];
#pragma unroll
    for (int c = 0; c < CB; c++) {
      const int k = k0 - c;
      const bool in = k >= 0;
      dv[c] = in ? a[k * NEQ + k] : 1.0;
      u0[c] = (in && has0 && l < k) ? a[k * NEQ + l] : 0.0;
      u1[c] = (in && has1 && r1 < k) ? a[k * NEQ + r1] : 0.0;
    }
#pragma unroll
    for (int c = 0; c < CB; c++) {
      const int k = k0 - c;
      if (k >= 0) {
        const double bk = __shfl_sync(FULLW, (k < 32) ? b0 : b1, k & 31) / dv[c];
        if (has0 && l < k) {
          b0 -= u0[c] * bk;
        }
        if (has1 && r1 < k) {
          b1 -= u1[c] * bk;
        }
        if (l == k) {
          b0 = bk;
        }
        if (r1 == k) {
          b1 = bk;
        }
      }
    }
  }
  __syncwarp();
  if (has0) {
    b[l] = b0;
  }
  if (has1) {
    b[r1] = b1;
  }
  __syncwarp();
}

template <int G>
PRX_MEM void BdfStoreCoop<G>::getrs(const double* a, const int* p, double* b) const
{
#ifndef PRX_COOP_GENERIC_LA
  if (G == 32 && NEQ > 32 && NEQ <= 64) {
    getrs_warp(a, perm(), b);
    return;
  }
#endif
  getrs_coop<G>(g, a, p, b);
}

// ---- thermo-chemistry, G lanes per cell -----------------------------------------------------------
// energy (ETYPE 1: internal energy, 2: enthalpy) and heat capacity of the mixture at T, in one pass;
// Y holds the mass fractions, each lane reads only the species it owns
template <class TB, int ETYPE, int G>
__device__ __forceinline__ void mix_e_cv_coop(const Grp<G>& g, const double T, const double* Y, double& e, double& cv)
{
  const TPow p = tpow_poly(T);
  double pe = 0.0, pc = 0.0;
#pragma unroll 1
  for (int k = g.gl; k < N; k += G) {
    const double yk = Y[k], imw = TB::imw(k);
    pe += yk * ((ETYPE == 1) ? sp_e_RT<TB>(k, p) : sp_h_RT<TB>(k, p)) * imw;
    pc += yk * ((ETYPE == 1) ? sp_cv_R<TB>(k, p) : sp_cp_R<TB>(k, p)) * imw;
  }
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) {
    pe += __shfl_xor_sync(g.gm, pe, o);
    pc += __shfl_xor_sync(g.gm, pc, o);
  }
  e = pe * (PRX_RU * T);
  cv = pc * PRX_RU;
}

// GET_T_GIVEN_EY / GET_T_GIVEN_HY (mechanism.H:1419-1520), same control flow; returns the heat capacity at
// the returned temperature in cv_out (what fKernelSpec evaluates next, ReactorUtils.H:389-392)
template <class TB, int ETYPE, int G>
__device__ __forceinline__ void t_from_energy_coop(const Grp<G>& g, const double e, const double* Y, double& t, double& cv_out)
{
  const double tmin = PRX_TMIN, tmax = PRX_TMAX;
  double emin, emax, cvmin, cvmax;
  mix_e_cv_coop<TB, ETYPE, G>(g, tmin, Y, emin, cvmin);
  mix_e_cv_coop<TB, ETYPE, G>(g, tmax, Y, emax, cvmax);
  bool cv_valid = false;
  double cv = 0.0;
  if (e < emin) {
    t = tmin - (emin - e) / cvmin;
  } else if (e > emax) {
    t = tmax - (emax - e) / cvmax;
  } else {
    double t1 = t;
    if (t1 < tmin || t1 > tmax) {
      t1 = tmin + (tmax - tmin) / (emax - emin) * (e - emin);
    }
#pragma unroll 1
    for (int i = 0; i < 200; ++i) {
      double e1;
      mix_e_cv_coop<TB, ETYPE, G>(g, t1, Y, e1, cv);
      cv_valid = true;
      double dt = (e - e1) / cv;
      if (dt > 100.) {
        dt = 100.;
      } else if (dt < -100.) {
        dt = -100.;
      } else if (fabs(dt) < 1.e-6) {
        break;
      }
      t1 += dt;
      cv_valid = false;
    }
    t = t1;
  }
  if (!cv_valid) {
    double e1;
    mix_e_cv_coop<TB, ETYPE, G>(g, t, Y, e1, cv);
  }
  cv_out = cv;
}

// utils::fKernelSpec (ReactorUtils.H:358-415) for one cell, G lanes.  y/ydot: NEQ vectors, sc/gs: N scratch,
// qv: reaction-rate scratch (all in the cell's shared-memory block).  rYsrc(n): external species source.
template <class TB, int G, class SrcF>
__device__ __forceinline__ void rhs_coop(
  const Grp<G>& g, const double t, const int reactor_type, const double* y, double* ydot, double* sc, double* gs,
  double* qv, const SrcScalars& c, SrcF rYsrc)
{
  g.sync();
  double part = 0.0;
#pragma unroll 1
  for (int n = g.gl; n < N; n += G) {
    part += y[n];
  }
  const double rho = g.sum(part);
  const double rho_inv = 1.0 / rho;
#pragma unroll 1
  for (int n = g.gl; n < N; n += G) {
    sc[n] = y[n] * rho_inv;
  }
  const double nrg = (c.rhoe_init + c.rhoesrc * t) * rho_inv;
  double temp = y[N];
  double Cv;
  if (reactor_type == 1) {
    t_from_energy_coop<TB, 1, G>(g, nrg, sc, temp, Cv);
  } else {
    t_from_energy_coop<TB, 2, G>(g, nrg, sc, temp, Cv);
  }
  const TPow p = tpow_full(temp);
  part = 0.0;
#pragma unroll 1
  for (int i = g.gl; i < N; i += G) {
    double C = rho * sc[i] * TB::imw(i);
    C *= 1.0e6;
    sc[i] = C;
    part += C;
    gs[i] = sp_gibbs<TB>(i, p);
  }
  const double mixture = g.sum(part);
  g.sync();
  const RxCommon rc = rx_common(p, mixture);
#pragma unroll 1
  for (int r = g.gl; r < PRX_NUM_REACTIONS; r += G) {
    qv[r] = reaction_q<TB>(r, rc, sc, gs);
  }
  g.sync();
  const double RT = PRX_RU * temp;
  part = 0.0;
#pragma unroll 1
  for (int n = g.gl; n < N; n += G) {
    double w = 0.0;
    const int j1 = TB::sp_ptr(n + 1);
    int j = TB::sp_ptr(n);
    // the gather list (reaction index, coefficient) and the rates are fetched four terms ahead of the sequential sum
    // (ncu r02h: the one-term-at-a-time chain table load -> shared-memory load -> FMA was 9 % of dodecane_lu's stall samples)
#pragma unroll 1
    for (; j + 4 <= j1; j += 4) {
      int rr[4];
      double cc[4], qq[4];
#pragma unroll
      for (int u = 0; u < 4; u++) {
        rr[u] = TB::sp_r(j + u);
        cc[u] = (double)TB::sp_c(j + u);
      }
#pragma unroll
      for (int u = 0; u < 4; u++) {
        qq[u] = qv[rr[u]];
      }
#pragma unroll
      for (int u = 0; u < 4; u++) {
        w = w + cc[u] * qq[u];
      }
    }
#pragma unroll 1
    for (; j < j1; j++) {
      w = w + (double)TB::sp_c(j) * qv[TB::sp_r(j)];
    }
    w *= 1.0e-6;
    w *= TB::mw(n);
    const double cdot_rYs = w + rYsrc(n);
    ydot[n] = cdot_rYs;
    double ei = (reactor_type == 1) ? sp_e_RT<TB>(n, p) : sp_h_RT<TB>(n, p);
    ei *= RT * TB::imw(n);
    part += cdot_rYs * ei;
  }
  const double tot = g.sum(part);
  if ((N & (G - 1)) == g.gl) {
    ydot[N] = (c.rhoesrc - tot) * (rho_inv / Cv);
  }
  g.sync();
}

// ---- analytic Jacobian, G lanes per cell ----------------------------------------------------------
// Phase 1 (lane-parallel over a chunk of G reactions): reaction_jac - rate constants, equilibrium, falloff
// and their T derivatives - results stay in the lane's registers.  Phase 2 (the chunk's reactions one after
// the other, in the reference's order): the five numbers of reaction t are broadcast from lane t by shuffle
// and the lanes scatter nu_m * dq/dC_k into the matrix (raw aJacobian, accumulated in the cell's Newton-matrix
// block of shared memory, which is free at this point of a linear setup).  Then the temperature row
// (shuffle reductions) and the DWDOT / cJac scaling into savedJ (global, L2 resident).  The result is the
// reference's aJacobian entry for entry (prx_tchem.cuh ajacobian_tab is the host-tested serial form).
template <class TB, int G>
__device__ __forceinline__ void jac_coop(
  const Grp<G>& g, const int reactor_type, const double* y, double* __restrict__ J, double* __restrict__ sc,
  double* __restrict__ gs, double* __restrict__ dq, double* __restrict__ wk, double* __restrict__ Jout)
{
  static_assert(PRX_NUM_REACTIONS >= 2 * N, "jac_coop keeps wdot[N] and h[N] in the reaction-rate scratch");
  const int consP = (reactor_type == 2);
  double* wdot = wk;
  double* h = wk + N;
  const int base_lane = (int)(threadIdx.x & 31u) & ~(G - 1);
  g.sync();
  double part = 0.0;
#pragma unroll 1
  for (int k = g.gl; k < N; k += G) {
    part += y[k];
  }
  const double rho = g.sum(part);
  const double temp = y[N];
  part = 0.0;
#pragma unroll 1
  for (int k = g.gl; k < N; k += G) {
    const double Yi = y[k] / rho;
    const double c = 1.e6 * (rho * Yi * TB::imw(k));
    sc[k] = c;
    part += c;
  }
  const double mixture = g.sum(part);
#pragma unroll 1
  for (int e = g.gl; e < NEQ * NEQ; e += G) {
    J[e] = 0.0;
  }
  const JacCommon c = jac_common(temp, mixture);
  TPow p = tpow_poly(temp);
  p.logT = c.logT;
#pragma unroll 1
  for (int k = g.gl; k < N; k += G) {
    gs[k] = sp_gibbs<TB>(k, p);
    h[k] = sp_h_RT<TB>(k, p);
    wdot[k] = 0.0;
    dq[k] = 0.0;
  }
  g.sync();
#pragma unroll 1
  for (int r0 = 0; r0 < PRX_NUM_REACTIONS; r0 += G) {
    RxJac x;
    x.q = x.dqdT = x.kf = x.kr = x.dcdc = 0.0;
    if (r0 + g.gl < PRX_NUM_REACTIONS) {
      x = reaction_jac<TB>(r0 + g.gl, c, sc, gs, h);
    }
    const int nrx = (PRX_NUM_REACTIONS - r0 < G) ? (PRX_NUM_REACTIONS - r0) : G;
#pragma unroll 1
    for (int t = 0; t < nrx; t++) {
      const int r = r0 + t;
      RxJac b;
      b.q = __shfl_sync(g.gm, x.q, base_lane + t);
      b.dqdT = __shfl_sync(g.gm, x.dqdT, base_lane + t);
      b.kf = __shfl_sync(g.gm, x.kf, base_lane + t);
      b.kr = __shfl_sync(g.gm, x.kr, base_lane + t);
      b.dcdc = __shfl_sync(g.gm, x.dcdc, base_lane + t);
      const int jm = TB::jmeta(r);
      const int nnu = jm & 15, nd = (jm >> 4) & 15;
      const bool row_lane = g.gl < nnu; // G >= 8 > 5 rows
      const int mrow = row_lane ? (int)TB::jnu_idx(5 * r + g.gl) : 0;
      const double crow = row_lane ? (double)TB::jnu_c(5 * r + g.gl) : 0.0;
      if (row_lane) {
        wdot[mrow] += crow * b.q;
      }
      if (rx_type(TB::meta(r)) == 0) {
#pragma unroll 1
        for (int d = 0; d < nd; d++) {
          const int k = TB::jd_k(6 * r + d);
          const double v = direct_dqdc<TB>(r, d, b, sc);
          if (row_lane) {
            J[k * NEQ + mrow] += crow * v;
          }
        }
      } else {
        const int tb = TB::tb_of(r);
#pragma unroll 1
        for (int k = g.gl; k < N; k += G) {
          int d = -1;
          for (int dd = 0; dd < nd; dd++) {
            if (TB::jd_k(6 * r + dd) == k) {
              d = dd;
            }
          }
          double v = 0.0;
          bool have = false;
          if (consP) {
            if (TB::tb_has(tb * N + k)) {
              have = true;
              v = (TB::tb_eff(tb * N + k)) * b.dcdc;
              if (d >= 0) {
                const int pf = TB::jd_f(6 * r + d), pr = TB::jd_r(6 * r + d);
                if (pf != 0) {
                  v = v + direct_term(pf, b.kf, sc);
                }
                if (pr != 0) {
                  v = v - direct_term(pr, b.kr, sc);
                }
              }
            } else if (d >= 0) {
              have = true;
              v = direct_dqdc<TB>(r, d, b, sc);
            }
          } else {
            // reference quirk: dqdc[k] keeps an earlier reaction's value when 1 + eff_k == 0 and k has no direct term
            const double tot = TB::tb_tot(tb * N + k);
            have = true;
            if (tot != 0.0) {
              v = (tot == 1.0) ? b.dcdc : (tot)*b.dcdc;
              if (d >= 0) {
                const int pf = TB::jd_f(6 * r + d), pr = TB::jd_r(6 * r + d);
                if (pf != 0) {
                  v = v + direct_term(pf, b.kf, sc);
                }
                if (pr != 0) {
                  v = v - direct_term(pr, b.kr, sc);
                }
              }
              dq[k] = v;
            } else if (d >= 0) {
              v = direct_dqdc<TB>(r, d, b, sc);
              dq[k] = v;
            } else {
              v = dq[k];
            }
          }
          if (have) {
            for (int j = 0; j < nnu; j++) {
              J[k * NEQ + TB::jnu_idx(5 * r + j)] += ((double)TB::jnu_c(5 * r + j)) * v;
            }
          }
        }
      }
      if (row_lane) {
        J[N * NEQ + mrow] += crow * b.dqdT;
      }
      g.sync();
    }
  }
  // temperature row; gs <- c_R, h <- e_RT (or h_RT)
  double p1 = 0.0, p2 = 0.0, p3 = 0.0, p4 = 0.0;
#pragma unroll 1
  for (int k = g.gl; k < N; k += G) {
    const double c_R = consP ? sp_cp_R<TB>(k, p) : sp_cv_R<TB>(k, p);
    const double eh = consP ? h[k] : sp_e_RT<TB>(k, p);
    const double dcR = sp_dcvpRdT<TB>(k, p);
    gs[k] = c_R;
    h[k] = eh;
    p1 += c_R * sc[k];
    p2 += dcR * sc[k];
    p3 += eh * wdot[k];
    p4 += c.invT * (c_R - eh) * wdot[k] + eh * J[N * NEQ + k];
  }
  const double cmix = g.sum(p1), dcmixdT = g.sum(p2), ehmix = g.sum(p3), dehmixdT = g.sum(p4);
  const double cmixinv = 1.0 / cmix;
  const double tmp1 = ehmix * cmixinv;
  const double tmp3 = cmixinv * temp;
  const double tmp2 = tmp1 * tmp3;
  g.sync();
#pragma unroll 1
  for (int k = g.gl; k < N; k += G) {
    double dehmixdc = 0.0;
#pragma unroll 1
    for (int mm = 0; mm < N; mm++) {
      dehmixdc += h[mm] * J[k * NEQ + mm];
    }
    J[k * NEQ + N] = tmp2 * gs[k] - tmp3 * dehmixdc;
  }
  if (g.gl == 0) {
    J[N * NEQ + N] = -tmp1 + tmp2 * dcmixdT - tmp3 * dehmixdT;
  }
  g.sync();
  // DWDOT unit conversion + cJac scaling between C and rhoY variables (ReactorCvodeJacobian.cpp:132-147)
#pragma unroll 1
  for (int e = g.gl; e < NEQ * NEQ; e += G) {
    const int col = e / NEQ;
    const int row = e - col * NEQ;
    double v = J[e];
    if (col < N && row < N) {
      v = v * TB::mw(row) * TB::imw(col);
    } else if (row == N && col < N) {
      v = (v * 1.e6) * TB::imw(col);
    } else if (col == N && row < N) {
      v = (v * 1.e-6) * TB::mw(row);
    }
    Jout[e] = v;
  }
  g.sync();
}

} // namespace prx
#endif
