// TEST INFRASTRUCTURE: host instantiation of the product's resumable integrators
// (pelec_b200/csrc/prx_integrators.cuh) so that the state-machine LOGIC can be unit-tested
// against the oracle on a machine without a GPU.  Never linked into libpelerx_b200.so.
#define PRX_HOST_TEST 1
#include <cstring>
#include "../../pelec_b200/csrc/prx_integrators.cuh"

using namespace prx;

static double* vsel(BdfCell& s, int sel)
{
  switch (sel) {
  case V_ZN0: return s.zn[0];
  case V_ZN1: return s.zn[1];
  case V_Y: return s.y;
  case V_TEMPV: return s.tempv;
  default: return s.ftemp;
  }
}

extern "C" {
int hh_num_species() { return N; }

// SoA box arrays a[comp*ncells + cell]; stats[cell*8 + {nst,nfe,nfeDQ,nje,nsetups,nni,ncfn,netf}]
int hh_react_cvode(
  int reactor_type, double rtol, double atol, const double* typ_vals, int max_order, int maxcor,
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
  static BdfCell s;
  for (int c = 0; c < ncells; c++) {
    if (mask && mask[c] == -1) {
      FC[c] = 0;
      continue;
    }
    memset(&s, 0, sizeof(s));
    for (int k = 0; k < N; k++) {
      s.zn[0][k] = rY[k * ncells + c];
      s.src.rYsrc[k] = rYsrc[k * ncells + c];
    }
    s.zn[0][N] = T[c];
    s.src.rhoe_init = rE[c];
    s.src.rhoesrc = rEsrc[c];
    bdf_begin(s, p, dt, time);
    for (;;) {
      const int req = bdf_resume(s, p);
      if (req == REQ_RHS) {
        rhs(s.t_eval, p.reactor_type, vsel(s, s.in_sel), vsel(s, s.out_sel), s.src);
      } else if (req == REQ_JAC) {
        double* A = s.A;
        jac_analytic(p.reactor_type, s.y, s.savedJ, [A](int row, int col, double v) { A[col * NEQ + row] = v; });
      } else if (req == REQ_LU) {
        s.lu_info = dense_getrf(s.A, s.piv);
      } else {
        break;
      }
    }
    // box_unflatten
    double rho = 0.0, Y[N];
    for (int k = 0; k < N; k++) {
      rY[k * ncells + c] = s.y[k];
      rho += s.y[k];
    }
    for (int k = 0; k < N; k++) {
      Y[k] = s.y[k] * (1.0 / rho);
    }
    rE[c] = s.src.rhoe_init + dt * s.src.rhoesrc;
    double temp = s.y[N];
    t_from_energy(reactor_type, rE[c] * (1.0 / rho), Y, temp);
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

int hh_react_rk64(
  int reactor_type, double atol, int nsg, int nsmin, int nsmax, int ncells, double* rY,
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
      rhs(s.t_eval, p.reactor_type, s.soln, s.ydot, s.src);
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
}
