// TEST INFRASTRUCTURE: host instantiation of the product's resumable integrators
// (pelec_b200/csrc/prx_integrators.cuh) so that the state-machine LOGIC can be unit-tested
// against the oracle on a machine without a GPU.  Never linked into libpelerx_b200.so.
#define PRX_HOST_TEST 1
#include <cstring>
#include "../../pelec_b200/csrc/prx_integrators.cuh"
#include "../../pelec_b200/csrc/prx_tchem.cuh"

#include <vector>

using namespace prx;

// One CVODE-per-cell integration of a box on the host.  Cell = BdfCellT<Store>; TAB selects the
// table-driven RHS (prx_tchem.cuh, tables TB) instead of the generated straight-line one.
template <class Cell, bool TAB, class TB>
static int react_cvode_host(
  Cell& s, int reactor_type, double rtol, double atol, const double* typ_vals, int max_order, int maxcor,
  int analytic_jac, int ncells, double* rY, const double* rYsrc, double* T, double* rE,
  const double* rEsrc, double* FC, const int* mask, double dt, double time, long* stats)
{
  BdfParams p;
  p.reactor_type = reactor_type;
  p.rtol = rtol;
  for (int k = 0; k < NEQ; k++) {
    p.atolv[k] = (typ_vals && typ_vals[0] > 0.0) ? typ_vals[k] * atol : atol;
  }
  p.qmax = max_order;
  p.maxcor = maxcor;
  p.analytic_jac = analytic_jac;
  p.mxstep = 10000;
  p.maxnef = 100;
  p.msbj = 100;
  p.clean_init_massfrac = 0;
  p.array4_semantics = 1;
  double scr[2 * N];
  for (int c = 0; c < ncells; c++) {
    if (mask && mask[c] == -1) {
      FC[c] = 0;
      continue;
    }
    for (int k = 0; k < N; k++) {
      s.zn(0)[k] = rY[k * ncells + c];
      s.src.rYsrc[k] = rYsrc[k * ncells + c];
    }
    s.zn(0)[N] = T[c];
    s.src.rhoe_init = rE[c];
    s.src.rhoesrc = rEsrc[c];
    bdf_begin(s, p, dt, time);
    for (;;) {
      const int req = bdf_resume(s, p);
      if (req == REQ_RHS) {
        if (TAB) {
          const SrcScalars sc{s.src.rhoe_init, s.src.rhoesrc};
          const double* src = s.src.rYsrc;
          rhs_tab<TB>(s.t_eval, p.reactor_type, s.vec(s.in_sel), s.vec(s.out_sel), scr, s.delta(), sc,
                      [src](int n) { return src[n]; });
        } else {
          double yin[NEQ], yout[NEQ];
          for (int k = 0; k < NEQ; k++) {
            yin[k] = s.vec(s.in_sel)[k];
          }
          rhs(s.t_eval, p.reactor_type, yin, yout, s.src);
          for (int k = 0; k < NEQ; k++) {
            s.vec(s.out_sel)[k] = yout[k];
          }
        }
      } else if (req == REQ_JAC) {
        double* J = s.savedJ(); // raw aJacobian output scaled in place (each entry is read once, then written)
        double yy[NEQ];
        for (int k = 0; k < NEQ; k++) {
          yy[k] = s.y()[k];
        }
        jac_analytic(p.reactor_type, yy, J, [J](int row, int col, double v) { J[col * NEQ + row] = v; });
      } else if (req == REQ_LU) {
        s.lu_info = dense_getrf(s.A(), s.piv());
      } else {
        break;
      }
    }
    // box_unflatten
    double rho = 0.0, Y[N];
    for (int k = 0; k < N; k++) {
      rY[k * ncells + c] = s.y()[k];
      rho += s.y()[k];
    }
    for (int k = 0; k < N; k++) {
      Y[k] = s.y()[k] * (1.0 / rho);
    }
    rE[c] = s.src.rhoe_init + dt * s.src.rhoesrc;
    double temp = s.y()[N];
    if (TAB) {
      t_from_energy_tab<TB>(reactor_type, rE[c] * (1.0 / rho), Y, temp);
    } else {
      t_from_energy(reactor_type, rE[c] * (1.0 / rho), Y, temp);
    }
    T[c] = temp;
    FC[c] = s.nfe + s.nfeDQ;
    if (stats) {
      long* st = stats + 8 * c;
      st[0] = s.nst; st[1] = s.nfe; st[2] = s.nfeDQ; st[3] = s.nje; st[4] = s.nsetups; st[5] = s.nni;
      st[6] = s.ncfn; st[7] = s.netf;
    }
  }
  return 0;
}

extern "C" {
int hh_num_species() { return N; }

// SoA box arrays a[comp*ncells + cell]; stats[cell*8 + {nst,nfe,nfeDQ,nje,nsetups,nni,ncfn,netf}]
// variant 0: thread-local arrays + generated RHS; 1: strided "shared-memory column" store with A in the
// column + table RHS (constant-space tables); 2: strided store, A local + table RHS (global-space tables)
int hh_react_cvode_variant(
  int variant, int reactor_type, double rtol, double atol, const double* typ_vals, int max_order, int maxcor,
  int analytic_jac, int ncells, double* rY, const double* rYsrc, double* T, double* rE,
  const double* rEsrc, double* FC, const int* mask, double dt, double time, long* stats)
{
  if (variant == 0) {
    static BdfCell s;
    memset(&s, 0, sizeof(s));
    return react_cvode_host<BdfCell, false, TabC>(
      s, reactor_type, rtol, atol, typ_vals, max_order, maxcor, analytic_jac, ncells, rY, rYsrc, T, rE, rEsrc, FC,
      mask, dt, time, stats);
  }
  if (variant == 3) {
    static BdfCell s;
    memset(&s, 0, sizeof(s));
    return react_cvode_host<BdfCell, true, TabC>(
      s, reactor_type, rtol, atol, typ_vals, max_order, maxcor, analytic_jac, ncells, rY, rYsrc, T, rE, rEsrc, FC,
      mask, dt, time, stats);
  }
  constexpr int STRIDE = 3;
  if (variant == 1) {
    typedef BdfCellT<BdfStoreSmem<STRIDE, true>> Cell;
    static Cell s;
    memset(&s, 0, sizeof(s));
    std::vector<double> col((size_t)Cell::doubles_per_thread(max_order) * STRIDE + 8, 0.0);
    s.base = col.data() + 1;
    s.nz = max_order + 1;
    return react_cvode_host<Cell, true, TabC>(
      s, reactor_type, rtol, atol, typ_vals, max_order, maxcor, analytic_jac, ncells, rY, rYsrc, T, rE, rEsrc, FC,
      mask, dt, time, stats);
  }
  typedef BdfCellT<BdfStoreSmem<STRIDE, false>> Cell;
  static Cell s;
  memset(&s, 0, sizeof(s));
  std::vector<double> col((size_t)Cell::doubles_per_thread(max_order) * STRIDE + 8, 0.0);
  s.base = col.data() + 2;
  s.nz = max_order + 1;
  return react_cvode_host<Cell, true, TabG>(
    s, reactor_type, rtol, atol, typ_vals, max_order, maxcor, analytic_jac, ncells, rY, rYsrc, T, rE, rEsrc, FC,
    mask, dt, time, stats);
}

int hh_react_cvode(
  int reactor_type, double rtol, double atol, const double* typ_vals, int max_order, int maxcor,
  int analytic_jac, int ncells, double* rY, const double* rYsrc, double* T, double* rE,
  const double* rEsrc, double* FC, const int* mask, double dt, double time, long* stats)
{
  return hh_react_cvode_variant(
    0, reactor_type, rtol, atol, typ_vals, max_order, maxcor, analytic_jac, ncells, rY, rYsrc, T, rE, rEsrc, FC, mask,
    dt, time, stats);
}

// n states, state-major: generated (tab = 0) or table-driven (tab = 1) RHS
int hh_rhs(int tab, int n, double t, int reactor_type, const double* y, const double* rhoe, const double* rhoesrc,
           const double* rYsrc, double* ydot)
{
  for (int s = 0; s < n; s++) {
    CellSrc src;
    for (int k = 0; k < N; k++) {
      src.rYsrc[k] = rYsrc[s * N + k];
    }
    src.rhoe_init = rhoe[s];
    src.rhoesrc = rhoesrc[s];
    if (tab) {
      double scr[2 * N];
      const SrcScalars sc{src.rhoe_init, src.rhoesrc};
      const double* sp = src.rYsrc;
      rhs_tab<TabG>(t, reactor_type, y + s * NEQ, ydot + s * NEQ, scr, scr + N, sc, [sp](int k) { return sp[k]; });
    } else {
      rhs(t, reactor_type, y + s * NEQ, ydot + s * NEQ, src);
    }
  }
  return 0;
}

// analytic Jacobian of n states (state-major y, J[s][col][row]): generated straight-line (tab = 0) or
// table-driven (tab = 1) aJacobian
int hh_jac(int tab, int n, int reactor_type, const double* y, double* J)
{
  static double Jt[NEQ * NEQ];
  for (int s = 0; s < n; s++) {
    double* Jo = J + (size_t)s * NEQ * NEQ;
    auto st = [Jo](int row, int col, double v) { Jo[col * NEQ + row] = v; };
    if (tab) {
      jac_analytic_tab<TabG>(reactor_type, y + s * NEQ, Jt, st);
    } else {
      jac_analytic(reactor_type, y + s * NEQ, Jt, st);
    }
  }
  return 0;
}

// LU factor + solve of n column-major NEQ x NEQ systems: reference-order loops (small = 0) or the
// register-staged device forms (small = 1); returns the sum of the GETRF return codes
int hh_lu(int small, int n, double* A, int* piv, double* b)
{
  int rc = 0;
  for (int s = 0; s < n; s++) {
    double* a = A + (size_t)s * NEQ * NEQ;
    int* p = piv + (size_t)s * NEQ;
    double* x = b + (size_t)s * NEQ;
    int perm[NEQ];
    int* pm = (s % 2) ? perm : nullptr; // both row-interchange forms of the staged solve
    const int r = (small == 2) ? dense_getrf_mid(a, p, pm) : (small ? dense_getrf_small(a, p, pm) : dense_getrf(a, p));
    rc += r;
    if (r == 0) {
      if (small) {
        dense_getrs_small(a, p, x, pm);
      } else {
        dense_getrs(a, p, x);
      }
    }
  }
  return rc;
}

int hh_react_rk64_variant(
  int tab, int reactor_type, double atol, int nsg, int nsmin, int nsmax, int ncells, double* rY,
  const double* rYsrc, double* T, double* rE, const double* rEsrc, double* FC, double dt, double time)
{
  Rk64Params p{reactor_type, atol, nsg, nsmin, nsmax, 1};
  static Rk64Cell s;
  for (int c = 0; c < ncells; c++) {
    memset(&s, 0, sizeof(s));
    for (int k = 0; k < N; k++) {
      s.soln[k] = rY[k * ncells + c];
      s.src.rYsrc[k] = rYsrc[k * ncells + c];
    }
    s.soln[N] = T[c];
    s.src.rhoe_init = rE[c];
    s.src.rhoesrc = rEsrc[c];
    rk64_begin(s, p, dt, time);
    while (rk64_resume(s, p) == REQ_RHS) {
      if (tab) {
        double scr[2 * N];
        const SrcScalars sc{s.src.rhoe_init, s.src.rhoesrc};
        const double* sp = s.src.rYsrc;
        rhs_tab<TabC>(s.t_eval, p.reactor_type, s.soln, s.ydot, scr, scr + N, sc, [sp](int k) { return sp[k]; });
      } else {
        rhs(s.t_eval, p.reactor_type, s.soln, s.ydot, s.src);
      }
    }
    rE[c] = rk64_finish(s, p);
    for (int k = 0; k < N; k++) {
      rY[k * ncells + c] = s.soln[k];
    }
    T[c] = s.soln[N];
    FC[c] = s.nsteps;
  }
  return 0;
}

int hh_react_rk64(
  int reactor_type, double atol, int nsg, int nsmin, int nsmax, int ncells, double* rY,
  const double* rYsrc, double* T, double* rE, const double* rEsrc, double* FC, double dt, double time)
{
  return hh_react_rk64_variant(0, reactor_type, atol, nsg, nsmin, nsmax, ncells, rY, rYsrc, T, rE, rEsrc, FC, dt, time);
}
}
