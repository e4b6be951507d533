// prx_kernels.cuh - shared declarations between the per-mechanism kernel translation units
// (prx_kernels.cu, compiled once per mechanism) and the mechanism-independent C ABI (prx_api.cu).
#ifndef PRX_KERNELS_CUH
#define PRX_KERNELS_CUH

#include <cuda_runtime.h>

#include "../../include/pelerx_b200.h"

namespace prx {

// One of the seven arrays crossing the reactor boundary, already reduced to what a kernel
// needs: base pointer of cell (lo) component 0, and strides in elements.
struct ArrView
{
  double* p;
  long long js, ks, ns;
};
struct ArrViewI
{
  const int* p;
  long long js, ks, ns;
};

// LAYOUT_STATE: the fused PeleC::react_state body (PC/Source/React.cpp:181-343): the kernel reads S_old / S_new /
// non_react_src itself (pack), integrates, and writes S_new and I_R (unpack + heat release) - no STemp / extsrc copies.
enum Layout { LAYOUT_ARRAY4 = 0, LAYOUT_FLAT = 1, LAYOUT_STATE = 2 };

// PeleC state-vector component indices (PC/Source/IndexDefines.H)
struct StateIx
{
  int urho, umx, umy, umz, ueden, ueint, utemp, ufs;
};

struct BoxArgs
{
  int layout;
  int ncells;
  int len[3];
  // Array4 layout
  ArrView rY, rYsrc, T, rE, rEsrc, FC;
  ArrViewI mask; // p == nullptr: no mask
  // flat layout (YCOrder): y[c*(N+1)+k], ysrc[c*N+k], e[c], esrc[c], fc[c]
  double* y;
  const double* ysrc;
  double* e;
  const double* esrc;
  double* fc;
  // state layout (fused react_state): S_old, S_new (in/out), non-reacting sources, I_R (N+2 components out); FC optional
  ArrView sold, snew, nonrs, IR;
  StateIx ix;
  int do_update; // !react_init
  // rotating frame (pelec.do_rf, React.cpp:123-131): rotenrg = 0.5 omega^2 r^2 with r the distance of the cell centre from
  // the axis; lo = global index of the box's first cell
  int rf_on, rf_axis, lo[3];
  double rf_omega, rf_axloc[3], rf_problo[3], rf_dx[3];
  double dt, time;
  const int* order; // optional: queue position -> cell index (cells sorted by temperature), nullptr = identity
};

// ---- pointwise EOS / transport kernels either side of the reactor ---------------------------------
// primitive-state component indices (PC/Source/IndexDefines.H: QRHO ... QFS, QGAMC ... QRSPEC)
struct PrimIx
{
  int qrho, qu, qv, qw, qgame, qpres, qreint, qtemp, qfs;
  int qgamc, qc, qcsml, qdpdr, qdpde, qrspec;
};
struct PrimArgs
{
  int ncells;
  int len[3];
  ArrView u, q, qa;             // ctoprim: conserved state in, primitives + auxiliaries out
  ArrView Y, T, rho;            // transport: inputs (component-offset views of q)
  ArrView D, mu, xi, lam;       // transport: outputs (D: N components)
  StateIx ix;
  PrimIx qx;
  double small_num;             // PeleC constants::small_num()
};

// ---- ReactorArkode (box-coupled explicit Runge-Kutta, prx_api.cu react_arkode): one stage = a linear combination of
// the stage derivatives + the ODE right-hand side of every cell; vectors are CYOrder columns v[k * ncells + c]
struct ArkStageArgs
{
  int ncells, reactor_type, nterms;
  double t;
  double coef[13];
  const double* F[13];
  const double* yn;
  double* ycur; // may be nullptr when nterms == 0 (f at yn itself)
  double* fout;
  const double* rhoe;
  const double* rhoesrc;
  const double* rYsrc;
};

// device-side counters (unsigned long long so that atomicAdd applies)
struct DevStats
{
  unsigned long long v[14];
};
enum StatIdx {
  ST_CELLS = 0, ST_MASKED, ST_STEPS, ST_RHS, ST_JAC, ST_SETUPS, ST_NEWTON, ST_ERRFAIL, ST_CONVFAIL,
  ST_FAILED, ST_MAXSTEPS, ST_LAST_STEPS,
  ST_LAST_MAXRHS // largest nfe + nfeDQ of a cell of the last react call (what the 1-D CVODE react returns)
};

// scheduling knobs of the warp loop (see prx_kernels.cu)
struct SchedParams
{
  int jac_defer_min_lanes; // serve REQ_JAC only when >= this many lanes ask (or nobody has RHS work)
  int lu_defer_min_lanes;
  int bdf_variant;      // BDF kernel generation (-1: the mechanism's default), see launch_bdf
  int max_warps_per_sm; // cap on resident warps of the shared-memory-column kernels (0: occupancy limit)
  int coop_a_global;    // cooperative kernel: Newton matrix in global memory (1), shared memory (0), by size (-1)
  int lu_defer_block;   // gen-1 kernel: percent of the block's threads that must hold a Jacobian/LU request before they are served (0: off)
  int lu_defer_max_age; // ... but never wait longer than this many iterations
  int refill_min;       // gen-1 kernel: idle lanes of a warp wait until this many are free before they take new cells
  int sort_cells;       // hand cells out sorted by temperature (1) or in index order (0)
  int block_warps;      // warps per block of the hot/cold kernel (1: independent warps; 2-4: phase-aligned by a block barrier)
};

struct MechOps
{
  int num_species;
  int num_reactions;
  const char* name;
  const char* const* species_names;
  // launchers; all asynchronous on `stream`
  cudaError_t (*launch_rk64)(const BoxArgs&, const prx_config&, int* queue, DevStats* stats, cudaStream_t stream);
  cudaError_t (*launch_bdf)(
    const BoxArgs&, const prx_config&, const double* typ_vals_host, int* queue, DevStats* stats,
    const SchedParams&, cudaStream_t stream);
  const char* (*kernel_name)(int reactor_kind, const SchedParams&);
  cudaError_t (*probe_rhs)(
    int n, double t, int reactor_type, const double* y, const double* rhoe, const double* rhoesrc,
    const double* rYsrc, double* ydot, cudaStream_t stream);
  cudaError_t (*probe_jac)(int n, int reactor_type, const double* y, double* J, cudaStream_t stream);
  cudaError_t (*flatten)(
    const BoxArgs&, const prx_config&, int ordering, double* y, double* ysrc, double* e, double* esrc, cudaStream_t stream);
  cudaError_t (*unflatten)(
    const BoxArgs&, const prx_config&, int ordering, const double* y, const double* e, const long long* fcunt,
    cudaStream_t stream);
  // unfused form of the react_state body for integrator kernels without a fused load/store phase (cooperative kernel):
  // pack fills the Array4 views of the same BoxArgs (rY, T, rE, rEsrc scratch; rYsrc aliases nonrs), unpack consumes them
  cudaError_t (*state_pack)(const BoxArgs&, cudaStream_t stream);
  cudaError_t (*state_unpack)(const BoxArgs&, const prx_config&, cudaStream_t stream);
  int (*bdf_is_fused)(const SchedParams&); // 1: launch_bdf handles LAYOUT_STATE itself
  cudaError_t (*compute_temp)(const BoxArgs&, cudaStream_t stream); // pc_cmpTemp on a.snew
  cudaError_t (*ctoprim)(const PrimArgs&, cudaStream_t stream);      // pc_ctoprim: u -> q, qa
  cudaError_t (*transport)(const PrimArgs&, cudaStream_t stream);    // SimpleTransport::get_transport_coeffs
  cudaError_t (*ark_stage)(const ArkStageArgs&, cudaStream_t stream); // ReactorArkode: stage state + cF_RHS
};

// one per kernel translation unit of this build (gen/mech_registry.inc, written by the Makefile from MECHS)
#define PRX_MECH_X(name) const MechOps* mech_ops_##name();
#include "gen/mech_registry.inc"
#undef PRX_MECH_X

} // namespace prx
#endif
