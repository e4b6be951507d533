// TEST INFRASTRUCTURE - not part of the product.
//
// A C door onto a LIVE, UNMODIFIED PeleC level (the PMF regression case compiled by oracle/ref_pelec/Makefile): start-up exactly
// as PC/Source/main.cpp:41-138 does it (amrex::Initialize, PeleCAmr, EB2, Amr::init), then read / write the level's state
// MultiFabs as flat arrays and call PeleC::react_state (PC/Source/React.cpp:30-362) on them.  Purpose: golden fixtures that PIN
// the react_state restatement (oracle/pyoracle.py react_state_oracle) and the fused prx_react_state_box kernels.
// Loaded with ctypes by tests/golden/make_golden3.py in this container only.
//
// Flat layout of a level: the valid cells of box 0, box 1, ... concatenated (i fastest within a box), component-major:
// a[n * ncells + c].
#include <memory>
#include <string>
#include <vector>

#include <AMReX.H>
#include <AMReX_Amr.H>
#include <AMReX_AmrLevel.H>
#include <AMReX_EB2.H>
#include <AMReX_Gpu.H>
#include <AMReX_MultiFab.H>
#include <AMReX_ParmParse.H>

#include "PeleC.H"
#include "PeleCAmr.H"
#include "Utilities.H"

std::string inputs_name; // PC/Source/main.cpp:18 (IO.cpp writes it into job_info)

void initialize_EB2(
  const amrex::Geometry& geom,
  const int eb_max_level,
  const int max_level,
  const int coarsening,
  const amrex::Vector<amrex::IntVect>& ref_ratio,
  const amrex::IntVect& max_grid_size);

amrex::LevelBld* getLevelBld();

namespace {
PeleCAmr* g_amr = nullptr;
amrex::Real g_stop_time = -1.0;

void
override_default_parameters()
{
  amrex::ParmParse pp("eb2");
  if (not pp.contains("geom_type")) {
    std::string geom_type("all_regular");
    pp.add("geom_type", geom_type);
  }
}

PeleC&
level(int lev)
{
  return dynamic_cast<PeleC&>(g_amr->getLevel(lev));
}

amrex::MultiFab&
which_mf(int lev, int which)
{
  PeleC& l = level(lev);
  switch (which) {
  case 0:
    return l.get_old_data(State_Type);
  case 1:
    return l.get_new_data(State_Type);
  default:
    return l.get_new_data(Reactions_Type);
  }
}

// CUDA build (Makefile.cuda): the level's data is device memory; the host loops below then run on a pinned-memory copy
// (host_copy / push_back), never on the level's own FABs
amrex::MultiFab
host_copy(const amrex::MultiFab& mf, bool with_data)
{
#ifdef AMREX_USE_GPU
  amrex::MultiFab h(mf.boxArray(), mf.DistributionMap(), mf.nComp(), 0, amrex::MFInfo().SetArena(amrex::The_Pinned_Arena()));
#else
  amrex::MultiFab h(mf.boxArray(), mf.DistributionMap(), mf.nComp(), 0);
#endif
  if (with_data) {
    amrex::MultiFab::Copy(h, mf, 0, 0, mf.nComp(), 0);
  }
#ifdef AMREX_USE_GPU
  amrex::Gpu::Device::synchronize();
#endif
  return h;
}

void
push_back(amrex::MultiFab& mf, const amrex::MultiFab& h)
{
  amrex::MultiFab::Copy(mf, h, 0, 0, mf.nComp(), 0);
#ifdef AMREX_USE_GPU
  amrex::Gpu::Device::synchronize();
#endif
}

template <class F>
void
for_cells(const amrex::MultiFab& mf, F&& f)
{
  long c = 0;
  for (amrex::MFIter mfi(mf); mfi.isValid(); ++mfi) {
    const amrex::Box& bx = mfi.validbox();
    const auto lo = amrex::lbound(bx);
    const auto hi = amrex::ubound(bx);
    for (int k = lo.z; k <= hi.z; ++k) {
      for (int j = lo.y; j <= hi.y; ++j) {
        for (int i = lo.x; i <= hi.x; ++i) {
          f(mfi, i, j, k, c++);
        }
      }
    }
  }
}
} // namespace

extern "C" {

// argv as for the PeleC executable: inputs file first, then name=value overrides.  Returns the finest level.
int
pelecref_init(int argc, char** argv)
{
  amrex::Initialize(argc, argv, true, MPI_COMM_WORLD, override_default_parameters);
  inputs_name = argv[1];
  amrex::Real strt_time = 0.0;
  amrex::ParmParse pp;
  pp.query("strt_time", strt_time);
  pp.query("stop_time", g_stop_time);
  g_amr = new PeleCAmr(getLevelBld());
  amrex::AmrLevel::SetEBSupportLevel(amrex::EBSupport::full);
  amrex::AmrLevel::SetEBMaxGrowCells(PeleC::numGrow() + 1, PeleC::numGrow() + 1, PeleC::numGrow() + 1);
  initialize_EB2(
    g_amr->Geom(PeleC::getEBMaxLevel()), PeleC::getEBMaxLevel(), g_amr->maxLevel(), PeleC::getEBCoarsening(),
    g_amr->refRatio(), g_amr->maxGridSize(g_amr->maxLevel()));
  g_amr->init(strt_time, g_stop_time);
  return g_amr->finestLevel();
}

int
pelecref_nvar()
{
  return NVAR;
}

int
pelecref_num_species()
{
  return NUM_SPECIES;
}

// index of the first species / of the named conserved components in the state (IndexDefines.H)
void
pelecref_indices(int* out)
{
  out[0] = URHO;
  out[1] = UMX;
  out[2] = UMY;
  out[3] = UMZ;
  out[4] = UEDEN;
  out[5] = UEINT;
  out[6] = UTEMP;
  out[7] = UFS;
}

long
pelecref_ncells(int lev)
{
  long n = 0;
  const amrex::MultiFab& mf = level(lev).get_new_data(State_Type);
  for (amrex::MFIter mfi(mf); mfi.isValid(); ++mfi) {
    n += mfi.validbox().numPts();
  }
  return n;
}

int
pelecref_nboxes(int lev)
{
  return static_cast<int>(level(lev).get_new_data(State_Type).boxArray().size());
}

// lo/hi of every box of the level: out[6 * b + {0..2}] = lo, {3..5} = hi
void
pelecref_boxes(int lev, int* out)
{
  const amrex::MultiFab& mf = level(lev).get_new_data(State_Type);
  int b = 0;
  for (amrex::MFIter mfi(mf); mfi.isValid(); ++mfi, ++b) {
    const amrex::Box& bx = mfi.validbox();
    for (int d = 0; d < 3; ++d) {
      out[6 * b + d] = bx.smallEnd(d);
      out[6 * b + 3 + d] = bx.bigEnd(d);
    }
  }
}

double
pelecref_time(int lev)
{
  return level(lev).get_state_data(State_Type).curTime();
}

// advance the whole hierarchy by n coarse time steps (Amr::coarseTimeStep: hydro, diffusion, reactions - the application)
double
pelecref_coarse_steps(int n)
{
  for (int s = 0; s < n; ++s) {
    g_amr->coarseTimeStep(g_stop_time);
  }
  return g_amr->cumTime();
}

double
pelecref_last_dt(int lev)
{
  return g_amr->dtLevel(lev);
}

// which: 0 = S_old, 1 = S_new, 2 = I_R (Reactions_Type)
int
pelecref_ncomp(int lev, int which)
{
  if (which == 0 && !level(lev).get_state_data(State_Type).hasOldData()) {
    level(lev).get_state_data(State_Type).allocOldData();
  }
  return which_mf(lev, which).nComp();
}

void
pelecref_get(int lev, int which, double* out)
{
  if (which == 0 && !level(lev).get_state_data(State_Type).hasOldData()) {
    level(lev).get_state_data(State_Type).allocOldData();
  }
  const amrex::MultiFab mf = host_copy(which_mf(lev, which), true);
  const long nc = pelecref_ncells(lev);
  const int ncomp = mf.nComp();
  for_cells(mf, [&](const amrex::MFIter& mfi, int i, int j, int k, long c) {
    auto const& a = mf.const_array(mfi);
    for (int n = 0; n < ncomp; ++n) {
      out[n * nc + c] = a(i, j, k, n);
    }
  });
}

void
pelecref_set(int lev, int which, const double* in)
{
  if (which == 0 && !level(lev).get_state_data(State_Type).hasOldData()) {
    level(lev).get_state_data(State_Type).allocOldData();
  }
  amrex::MultiFab& dst = which_mf(lev, which);
  amrex::MultiFab mf = host_copy(dst, false);
  const long nc = pelecref_ncells(lev);
  const int ncomp = mf.nComp();
  for_cells(mf, [&](const amrex::MFIter& mfi, int i, int j, int k, long c) {
    auto const& a = mf.array(mfi);
    for (int n = 0; n < ncomp; ++n) {
      a(i, j, k, n) = in[n * nc + c];
    }
  });
  push_back(dst, mf);
}

// PeleC::react_state(time, dt, react_init, &aux_src) on level lev; aux_src [NVAR][ncells] = the non-reacting source the MOL /
// SDC drivers pass (React.cpp:83-87).  react_init != 0: the start-up call (no update of S_new, zero source; aux ignored).
void
pelecref_react_state(int lev, double dt, int react_init, const double* aux)
{
  PeleC& l = level(lev);
  amrex::MultiFab& S_new = l.get_new_data(State_Type);
  amrex::MultiFab src(S_new.boxArray(), S_new.DistributionMap(), NVAR, S_new.nGrow(), amrex::MFInfo(), l.Factory());
  src.setVal(0.0);
  if (aux != nullptr) {
    const long nc = pelecref_ncells(lev);
    amrex::MultiFab hsrc = host_copy(src, false);
    for_cells(hsrc, [&](const amrex::MFIter& mfi, int i, int j, int k, long c) {
      auto const& a = hsrc.array(mfi);
      for (int n = 0; n < NVAR; ++n) {
        a(i, j, k, n) = aux[n * nc + c];
      }
    });
    push_back(src, hsrc);
  }
  l.react_state(pelecref_time(lev), dt, react_init != 0, react_init != 0 ? nullptr : &src);
}

// pc_ctoprim (PC/Source/Utilities.H:242-321) on n cells: u [NVAR][n] -> q [QVAR][n], qa [NQAUX][n].  Host build only.
int
pelecref_qvar()
{
  return QVAR;
}

int
pelecref_nqaux()
{
  return NQAUX;
}

#ifndef AMREX_USE_GPU
void
pelecref_ctoprim(int n, const double* u, double* q, double* qa)
{
  const amrex::Box bx(amrex::IntVect(0, 0, 0), amrex::IntVect(n - 1, 0, 0));
  amrex::FArrayBox ufab(bx, NVAR), qfab(bx, QVAR), qafab(bx, NQAUX);
  auto const& ua = ufab.array();
  auto const& qq = qfab.array();
  auto const& qaa = qafab.array();
  for (int c = 0; c < NVAR; ++c) {
    for (int i = 0; i < n; ++i) {
      ua(i, 0, 0, c) = u[static_cast<long>(c) * n + i];
    }
  }
  qfab.setVal<amrex::RunOn::Host>(0.0);
  qafab.setVal<amrex::RunOn::Host>(0.0);
  for (int i = 0; i < n; ++i) {
    pc_ctoprim(i, 0, 0, ufab.const_array(), qq, qaa);
  }
  for (int c = 0; c < QVAR; ++c) {
    for (int i = 0; i < n; ++i) {
      q[static_cast<long>(c) * n + i] = qq(i, 0, 0, c);
    }
  }
  for (int c = 0; c < NQAUX; ++c) {
    for (int i = 0; i < n; ++i) {
      qa[static_cast<long>(c) * n + i] = qaa(i, 0, 0, c);
    }
  }
}
#endif

} // extern "C"
