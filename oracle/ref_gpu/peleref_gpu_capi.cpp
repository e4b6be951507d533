// TEST / MEASUREMENT INFRASTRUCTURE - not part of the product.
//
// Thin extern "C" door onto the UNMODIFIED reference built for CUDA (oracle/ref_gpu/Makefile): PelePhysics' own GPU
// reactors - ReactorCvode::react, AMREX_USE_GPU region (PP/Source/Reactions/ReactorCvode.cpp:1315-1376: the whole box is ONE
// coupled CVODE system on N_Vector_Cuda, GMRES / cuSolver batched QR linear solves) and ReactorRK64::react
// (ReactorRK64.cpp:166-320 as an amrex::ParallelFor kernel) - so that bench.py can state the reference's OWN GPU path on
// the same B200 beside this repo's kernels.  One shared object per mechanism:
// oracle/_ref/libpeleref_gpu_<mech>.so, loaded with ctypes by tools/ref_gpu_bench.py in its own process.
//
// The caller passes HOST arrays in FAB layout (component stride nx*ny*nz, i fastest); this file owns the device copies, calls
// ReactorBase::react(box, Array4..., stream) on them and times that call alone with CUDA events on the AMReX stream.
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include <AMReX.H>
#include <AMReX_Array4.H>
#include <AMReX_Box.H>
#include <AMReX_Gpu.H>
#include <AMReX_ParmParse.H>
#include <AMReX_Sundials.H>

#include "mechanism.H"
#include <PelePhysics.H>
#include <ReactorBase.H>

using amrex::Real;
using pele::physics::reactions::ReactorBase;

namespace {
bool g_init = false;

struct RefReactor
{
  std::unique_ptr<ReactorBase> r;
};

template <class T>
struct DevBuf
{
  T* p = nullptr;
  size_t n = 0;
  explicit DevBuf(size_t count) : n(count) { cudaMalloc(&p, n * sizeof(T)); }
  ~DevBuf() { cudaFree(p); }
  void up(const T* h) const { cudaMemcpy(p, h, n * sizeof(T), cudaMemcpyHostToDevice); }
  void down(T* h) const { cudaMemcpy(h, p, n * sizeof(T), cudaMemcpyDeviceToHost); }
};
} // namespace

extern "C" {

int
peleref_gpu_init()
{
  if (g_init) {
    return 0;
  }
  int argc = 1;
  static char arg0[] = "peleref_gpu";
  static char* argv_store[] = {arg0, nullptr};
  char** argv = argv_store;
  amrex::Initialize(argc, argv, false, MPI_COMM_WORLD, []() {
    amrex::ParmParse pp("amrex");
    pp.add("verbose", 0);
    pp.add("v", 0);
    pp.add("signal_handling", 0);
    pp.add("the_arena_is_managed", 0);
    pp.add("the_arena_init_size", 0); // grow on demand: the caller's process may hold GPU memory too
    pp.add("abort_on_out_of_gpu_memory", 1);
  });
#ifndef AMREX_USE_SUNDIALS
  amrex::sundials::Initialize(1);
#endif
  g_init = true;
  return 0;
}

int
peleref_gpu_num_species()
{
  return NUM_SPECIES;
}

void
peleref_gpu_set_parm(const char* prefix, const char* name, const char* value)
{
  amrex::ParmParse pp(prefix);
  if (pp.contains(name)) {
    pp.remove(name);
  }
  pp.add(name, std::string(value));
}

void*
peleref_gpu_reactor_create(const char* key, int reactor_type, int ncells)
{
  auto* h = new RefReactor;
  h->r = ReactorBase::create(std::string(key));
  h->r->init(reactor_type, ncells);
  return h;
}

void
peleref_gpu_reactor_destroy(void* hv)
{
  auto* h = static_cast<RefReactor*>(hv);
  h->r->close();
  delete h;
}

// nrep timed ReactorBase::react calls on an nx*ny*nz box, each from a fresh device copy of the caller's state; the arrays
// of the LAST call are copied back.  ms_out[r] = CUDA-event time of call r (the react call only).  Returns react's value.
int
peleref_gpu_react_box_host(
  void* hv,
  int nx,
  int ny,
  int nz,
  double* rY,
  const double* rYsrc,
  double* T,
  double* rE,
  const double* rEsrc,
  double* FC,
  const int* mask,
  double dt,
  double time,
  int nrep,
  float* ms_out)
{
  auto* h = static_cast<RefReactor*>(hv);
  const size_t nc = static_cast<size_t>(nx) * ny * nz;
  amrex::Box box(amrex::IntVect(0, 0, 0), amrex::IntVect(nx - 1, ny - 1, nz - 1));
  DevBuf<Real> d_rY(nc * NUM_SPECIES), d_rYsrc(nc * NUM_SPECIES), d_T(nc), d_rE(nc), d_rEsrc(nc), d_FC(nc);
  DevBuf<int> d_mask(nc);
  d_rYsrc.up(rYsrc);
  d_rEsrc.up(rEsrc);
  d_mask.up(mask);
  const amrex::Dim3 lo{0, 0, 0}, hi{nx, ny, nz};
  amrex::Array4<Real> a_rY(d_rY.p, lo, hi, NUM_SPECIES), a_rYsrc(d_rYsrc.p, lo, hi, NUM_SPECIES), a_T(d_T.p, lo, hi, 1),
    a_rE(d_rE.p, lo, hi, 1), a_rEsrc(d_rEsrc.p, lo, hi, 1), a_FC(d_FC.p, lo, hi, 1);
  amrex::Array4<int> a_mask(d_mask.p, lo, hi, 1);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  cudaStream_t st = amrex::Gpu::gpuStream();
  int rc = 0;
  for (int r = 0; r < nrep; ++r) {
    d_rY.up(rY);
    d_T.up(T);
    d_rE.up(rE);
    cudaMemset(d_FC.p, 0, nc * sizeof(Real));
    cudaDeviceSynchronize();
    Real dt_react = dt;
    Real t = time;
    cudaEventRecord(e0, st);
    rc = h->r->react(box, a_rY, a_rYsrc, a_T, a_rE, a_rEsrc, a_FC, a_mask, dt_react, t, st);
    cudaEventRecord(e1, st);
    cudaEventSynchronize(e1);
    cudaDeviceSynchronize();
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    if (ms_out != nullptr) {
      ms_out[r] = ms;
    }
  }
  d_rY.down(rY);
  d_T.down(T);
  d_rE.down(rE);
  d_FC.down(FC);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  return rc;
}

} // extern "C"
