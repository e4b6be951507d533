// TEST INFRASTRUCTURE — not part of the product.
//
// Thin extern "C" door onto the UNMODIFIED reference (PelePhysics + AMReX + SUNDIALS,
// compiled from the sources where they lie under /root/reference by oracle/ref/Makefile).
// One shared object per mechanism (NUM_SPECIES is a compile-time constant of the
// reference): oracle/_ref/libpeleref_<mech>.so.  It is loaded with ctypes by tests/,
// by __graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference legs to
//   (1) pin the C restatement in oracle/ against the real thing, and
//   (2) time the reference's own CPU reactor on the host cores.
//
// Entry points call straight into:
//   ReactorBase::create / init / set_typ_vals_ode / react  (PP/Source/Reactions/ReactorBase.H:14-106)
//   mechanism functions (PP/Mechanisms/<mech>/mechanism.H) and the Fuego EOS facade
//   (PP/Source/Eos/Fuego.H), utils::fKernelSpec (PP/Source/Reactions/ReactorUtils.H:358-415).
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include <AMReX.H>
#include <AMReX_Array4.H>
#include <AMReX_Box.H>
#include <AMReX_ParmParse.H>
#include <AMReX_Sundials.H>

#include "mechanism.H"
#include <PelePhysics.H>
#include <ReactorBase.H>
#include <Transport.H>

using amrex::Real;
using pele::physics::reactions::ReactorBase;

namespace {
bool g_init = false;

struct RefReactor
{
  std::unique_ptr<ReactorBase> r;
};

amrex::Array4<Real>
view(Real* p, int nx, int ny, int nz, int nc)
{
  return amrex::Array4<Real>(
    p, amrex::Dim3{0, 0, 0}, amrex::Dim3{nx, ny, nz}, nc);
}
} // namespace

extern "C" {

int
peleref_init()
{
  if (g_init) {
    return 0;
  }
  int argc = 1;
  static char arg0[] = "peleref";
  static char* argv_store[] = {arg0, nullptr};
  char** argv = argv_store;
  // no inputs file, no command-line parsing, quiet
  amrex::Initialize(
    argc, argv, false, MPI_COMM_WORLD, []() {
      amrex::ParmParse pp("amrex");
      pp.add("verbose", 0);
      pp.add("v", 0);
      pp.add("signal_handling", 0);
      pp.add("the_arena_init_size", 0);
    });
#ifndef AMREX_USE_SUNDIALS
  amrex::sundials::Initialize(1);
#endif
  g_init = true;
  return 0;
}

int
peleref_num_species()
{
  return NUM_SPECIES;
}

int
peleref_num_reactions()
{
  return NUM_REACTIONS;
}

// space-separated species names written into buf
int
peleref_species_names(char* buf, int buflen)
{
  amrex::Vector<std::string> names;
  CKSYMS_STR(names);
  std::string s;
  for (auto const& n : names) {
    s += n + " ";
  }
  std::strncpy(buf, s.c_str(), buflen - 1);
  buf[buflen - 1] = 0;
  return static_cast<int>(names.size());
}

void
peleref_set_parm(const char* prefix, const char* name, const char* value)
{
  amrex::ParmParse pp(prefix);
  if (pp.contains(name)) {
    pp.remove(name);
  }
  pp.add(name, std::string(value));
}

void*
peleref_reactor_create(const char* key, int reactor_type, int ncells)
{
  auto* h = new RefReactor;
  h->r = ReactorBase::create(std::string(key));
  h->r->init(reactor_type, ncells);
  return h;
}

void
peleref_reactor_set_typ_vals(void* hv, const double* typ, int n)
{
  auto* h = static_cast<RefReactor*>(hv);
  std::vector<Real> v(typ, typ + n);
  h->r->set_typ_vals_ode(v);
}

void
peleref_reactor_destroy(void* hv)
{
  auto* h = static_cast<RefReactor*>(hv);
  h->r->close();
  delete h;
}

// One ReactorBase::react(box, Array4...) call on an nx*ny*nz box held as SoA arrays
// (component stride = nx*ny*nz, i fastest), exactly the layout of an AMReX FAB.
int
peleref_react_box(
  void* hv,
  int nx,
  int ny,
  int nz,
  double* rY,          // [N][nz][ny][nx]  in/out
  double* rYsrc,       // [N][...]         in
  double* T,           // [1]              in/out
  double* rE,          // [1]              in/out
  double* rEsrc,       // [1]              in
  double* FC,          // [1]              out
  int* mask,           // [1]              in
  double dt,
  double time)
{
  auto* h = static_cast<RefReactor*>(hv);
  amrex::Box box(amrex::IntVect(0, 0, 0), amrex::IntVect(nx - 1, ny - 1, nz - 1));
  auto a_rY = view(rY, nx, ny, nz, NUM_SPECIES);
  auto a_rYsrc = view(rYsrc, nx, ny, nz, NUM_SPECIES);
  auto a_T = view(T, nx, ny, nz, 1);
  auto a_rE = view(rE, nx, ny, nz, 1);
  auto a_rEsrc = view(rEsrc, nx, ny, nz, 1);
  auto a_FC = view(FC, nx, ny, nz, 1);
  amrex::Array4<int> a_mask(
    mask, amrex::Dim3{0, 0, 0}, amrex::Dim3{nx, ny, nz}, 1);
  Real dt_react = dt;
  Real t = time;
  return h->r->react(
    box, a_rY, a_rYsrc, a_T, a_rE, a_rEsrc, a_FC, a_mask, dt_react, t);
}

// Flat (1-D, YCOrder) ReactorBase::react: y[ncells*(N+1)] incl. T, ysrc[ncells*N].
int
peleref_react_flat(
  void* hv,
  double* y,
  double* ysrc,
  double* rX,
  double* rXsrc,
  double dt,
  double time,
  int ncells)
{
  auto* h = static_cast<RefReactor*>(hv);
  Real dt_react = dt;
  Real t = time;
  return h->r->react(y, ysrc, rX, rXsrc, dt_react, t, ncells);
}

// ---------------------------------------------------------------- unit-level probes
// All take n states; arrays are state-major: Y[n][N], C[n][N] ...

void
peleref_pyt2re(int n, const double* P, const double* Y, const double* T, double* rho, double* e)
{
  auto eos = pele::physics::PhysicsType::eos();
  for (int s = 0; s < n; ++s) {
    Real yy[NUM_SPECIES];
    for (int k = 0; k < NUM_SPECIES; ++k) {
      yy[k] = Y[s * NUM_SPECIES + k];
    }
    Real r, en;
    eos.PYT2RE(P[s], yy, T[s], r, en);
    rho[s] = r;
    e[s] = en;
  }
}

void
peleref_x2y(int n, const double* X, double* Y)
{
  auto eos = pele::physics::PhysicsType::eos();
  for (int s = 0; s < n; ++s) {
    Real xx[NUM_SPECIES], yy[NUM_SPECIES];
    for (int k = 0; k < NUM_SPECIES; ++k) {
      xx[k] = X[s * NUM_SPECIES + k];
    }
    eos.X2Y(xx, yy);
    for (int k = 0; k < NUM_SPECIES; ++k) {
      Y[s * NUM_SPECIES + k] = yy[k];
    }
  }
}

// T from (e, Y) with the incoming T as the Newton guess: Fuego::REY2T -> GET_T_GIVEN_EY
void
peleref_ey2t(int n, const double* e, const double* Y, double* T)
{
  auto eos = pele::physics::PhysicsType::eos();
  for (int s = 0; s < n; ++s) {
    Real yy[NUM_SPECIES];
    for (int k = 0; k < NUM_SPECIES; ++k) {
      yy[k] = Y[s * NUM_SPECIES + k];
    }
    Real t = T[s];
    eos.REY2T(1.0, e[s], yy, t);
    T[s] = t;
  }
}

// thermo: ei[n][N] (CKUMS, erg/g), cv[n] (CKCVBS), e_mix[n] (CKUBMS), hi[n][N] (CKHMS)
void
peleref_thermo(int n, const double* T, const double* Y, double* ei, double* cv, double* emix, double* hi)
{
  for (int s = 0; s < n; ++s) {
    Real yy[NUM_SPECIES], out[NUM_SPECIES];
    for (int k = 0; k < NUM_SPECIES; ++k) {
      yy[k] = Y[s * NUM_SPECIES + k];
    }
    CKUMS(T[s], out);
    for (int k = 0; k < NUM_SPECIES; ++k) {
      ei[s * NUM_SPECIES + k] = out[k];
    }
    CKHMS(T[s], out);
    for (int k = 0; k < NUM_SPECIES; ++k) {
      hi[s * NUM_SPECIES + k] = out[k];
    }
    Real c, u;
    CKCVBS(T[s], yy, c);
    CKUBMS(T[s], yy, u);
    cv[s] = c;
    emix[s] = u;
  }
}

// wdot[n][N] in CGS mol/cm^3/s from C[n][N] (mol/cm^3): CKWC
void
peleref_ckwc(int n, const double* T, const double* C, double* wdot)
{
  for (int s = 0; s < n; ++s) {
    Real cc[NUM_SPECIES], w[NUM_SPECIES];
    for (int k = 0; k < NUM_SPECIES; ++k) {
      cc[k] = C[s * NUM_SPECIES + k];
    }
    CKWC(T[s], cc, w);
    for (int k = 0; k < NUM_SPECIES; ++k) {
      wdot[s * NUM_SPECIES + k] = w[k];
    }
  }
}

// Fuego::RTY2WDOT: mass production rates [g/cm^3/s]
void
peleref_rty2wdot(int n, const double* rho, const double* T, const double* Y, double* wdot)
{
  auto eos = pele::physics::PhysicsType::eos();
  for (int s = 0; s < n; ++s) {
    Real yy[NUM_SPECIES], w[NUM_SPECIES];
    for (int k = 0; k < NUM_SPECIES; ++k) {
      yy[k] = Y[s * NUM_SPECIES + k];
    }
    eos.RTY2WDOT(rho[s], T[s], yy, w);
    for (int k = 0; k < NUM_SPECIES; ++k) {
      wdot[s * NUM_SPECIES + k] = w[k];
    }
  }
}

// Fuego::RTY2JAC -> DWDOT: J[n][(N+1)^2], column-major, concentration variables (CGS)
void
peleref_rty2jac(int n, const double* rho, const double* T, const double* Y, double* J, int HP)
{
  auto eos = pele::physics::PhysicsType::eos();
  const int neq = NUM_SPECIES + 1;
  for (int s = 0; s < n; ++s) {
    Real yy[NUM_SPECIES];
    for (int k = 0; k < NUM_SPECIES; ++k) {
      yy[k] = Y[s * NUM_SPECIES + k];
    }
    Real jj[(NUM_SPECIES + 1) * (NUM_SPECIES + 1)];
    eos.RTY2JAC(rho[s], T[s], yy, jj, HP);
    for (int q = 0; q < neq * neq; ++q) {
      J[s * neq * neq + q] = jj[q];
    }
  }
}

// The shared ODE right-hand side: y[n][N+1] (rhoY, T) -> ydot[n][N+1]
void
peleref_fkernelspec(
  int n,
  double t,
  int reactor_type,
  const double* y,
  const double* rhoe_init,
  const double* rhoesrc_ext,
  const double* rYsrc,
  double* ydot)
{
  const int neq = NUM_SPECIES + 1;
  for (int s = 0; s < n; ++s) {
    pele::physics::reactions::utils::fKernelSpec<
      pele::physics::reactions::utils::YCOrder>(
      0, 1, t, reactor_type, y + s * neq, ydot + s * neq, rhoe_init + s,
      rhoesrc_ext + s, rYsrc + s * NUM_SPECIES);
  }
}

// SimpleTransport::transport as PeleC's diffusion driver calls it (PC/Source/Diffusion.cpp:195-222 through
// PC/Source/TransCoeff.H:9-41: xi, mu, lambda, Ddiag; no Soret): Y[n][N] -> D[n][N], mu[n], xi[n], lam[n]
void
peleref_transport(
  int n, const double* T, const double* rho, const double* Y, double* D, double* mu, double* xi, double* lam)
{
  static pele::physics::PeleParams<pele::physics::transport::TransParm<
    pele::physics::PhysicsType::eos_type, pele::physics::PhysicsType::transport_type>>
    trans_parms;
  static bool ready = false;
  if (!ready) {
    trans_parms.initialize();
    ready = true;
  }
  auto const* tparm = trans_parms.device_parm();
  auto trans = pele::physics::PhysicsType::transport();
  for (int s = 0; s < n; ++s) {
    Real yy[NUM_SPECIES], dd[NUM_SPECIES];
    for (int k = 0; k < NUM_SPECIES; ++k) {
      yy[k] = Y[s * NUM_SPECIES + k];
    }
    Real m = 0.0, x = 0.0, l = 0.0;
    trans.transport(true, true, true, true, false, T[s], rho[s], yy, dd, nullptr, m, x, l, tparm);
    for (int k = 0; k < NUM_SPECIES; ++k) {
      D[s * NUM_SPECIES + k] = dd[k];
    }
    mu[s] = m;
    xi[s] = x;
    lam[s] = l;
  }
}

void
peleref_mw(double* mw_out)
{
  get_mw(mw_out);
}

} // extern "C"
