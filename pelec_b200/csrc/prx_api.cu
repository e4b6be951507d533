// prx_api.cu - the C ABI of include/pelerx_b200.h: handle life cycle, argument checking,
// Array4 -> kernel views, per-mechanism dispatch, statistics.  No chemistry in this file and no
// CPU compute path: if there is no CUDA device, prx_create fails.
#include <cstdio>
#include <cstdlib>
#include <atomic>
#include <cmath>
#include <cstring>
#include <mutex>
#include <new>
#include <vector>

#include "prx_kernels.cuh"
#include "prx_arkode.cuh"

#include <cub/device/device_radix_sort.cuh>
#include <nvtx3/nvToolsExt.h> // header-only (dlopens the injection library of an attached profiler; no link dependency)

using namespace prx;

struct prx_handle
{
  prx_config cfg;
  const MechOps* ops;
  int device;
  std::vector<double> typ_vals; // size N+1; typ_vals[0] <= 0: not set
  DevStats* d_stats;
  cudaStream_t box_streams[8];   // internal streams of the multi-box entry points (created on first use)
  cudaEvent_t box_events[9];
  int n_box_streams;
  std::mutex host_mutex;        // the *_host calls share one staging buffer: serialised per handle
  SchedParams sched;
  long long last_ncells;
  int last_layout;
  // staging for the *_host convenience calls
  double* d_stage;
  size_t d_stage_bytes;
  cudaStream_t stage_stream;
  cudaStream_t in_stream, out_stream; // host staging pipeline of prx_react_box_host: H2D | react | D2H per slab
  cudaStream_t host_react_streams[3] = {nullptr, nullptr, nullptr}; // extra react streams of that pipeline (slab s reacts on stream s mod n: the drain of one slab's persistent kernel overlaps the next slab's start)
  bool keep_last_steps;              // react_args: do not reset the per-call step sum (slabs of one call)
  ark::Controller ark_ctl;           // ReactorArkode: the step-size controller's history persists across react calls
  ark::Table ark_tab;                //                the Butcher table of cfg.ark_method
  long long ark_last_nfe = 0, ark_nfe = 0, ark_nst = 0, ark_netf = 0, ark_too_much_work = 0, ark_cells = 0;
  int rf_on = 0, rf_axis = 2;        // rotating frame of prx_react_state_box(es) (prx_set_rotating_frame)
  double rf_omega = 0.0, rf_axloc[3] = {0, 0, 0}, rf_problo[3] = {0, 0, 0}, rf_dx[3] = {1, 1, 1};
  char err[512];
};

namespace {

// the mechanisms of this build in id order: LiDryer 0, drm19 1, dodecane_lu 2 (ABI constants), then the extra ones
typedef const MechOps* (*MechFn)();
#define PRX_MECH_X(name) mech_ops_##name,
const MechFn g_mech_table[] = {
#include "gen/mech_registry.inc"
};
#undef PRX_MECH_X
constexpr int g_num_mechs = (int)(sizeof(g_mech_table) / sizeof(g_mech_table[0]));

const MechOps* ops_for(int mech)
{
  return (mech >= 0 && mech < g_num_mechs) ? g_mech_table[mech]() : nullptr;
}

int fail(prx_handle* h, int code, const char* what, cudaError_t ce = cudaSuccess)
{
  if (h != nullptr) {
    if (ce != cudaSuccess) {
      std::snprintf(h->err, sizeof(h->err), "%s: %s", what, cudaGetErrorString(ce));
    } else {
      std::snprintf(h->err, sizeof(h->err), "%s", what);
    }
  }
  return code;
}

#define PRX_CUDA(h, call)                            \
  do {                                               \
    const cudaError_t ce__ = (call);                 \
    if (ce__ != cudaSuccess) {                       \
      return fail((h), PRX_ERR_CUDA, #call, ce__);   \
    }                                                \
  } while (0)

struct DeviceGuard
{
  int prev;
  bool active;
  explicit DeviceGuard(int dev) : prev(0), active(false)
  {
    if (cudaGetDevice(&prev) == cudaSuccess && prev != dev) {
      active = (cudaSetDevice(dev) == cudaSuccess);
    }
  }
  ~DeviceGuard()
  {
    if (active) {
      cudaSetDevice(prev);
    }
  }
};

bool view_of(const prx_array4* a, const int lo[3], int need_comp, ArrView* v, const int* hi = nullptr)
{
  if (a == nullptr || a->p == nullptr || a->ncomp < need_comp) {
    return false;
  }
  for (int d = 0; d < 3; d++) {
    if (lo[d] < a->begin[d] || (hi != nullptr && hi[d] >= a->end[d])) { // the box must lie inside the array (end is exclusive)
      return false;
    }
  }
  v->js = a->jstride;
  v->ks = a->kstride;
  v->ns = a->nstride;
  v->p = a->p + (lo[0] - a->begin[0]) + (long long)(lo[1] - a->begin[1]) * a->jstride +
         (long long)(lo[2] - a->begin[2]) * a->kstride;
  return true;
}

// One work-queue counter per launch, from the stream-ordered pool of the launch's stream: released (cudaFreeAsync) behind
// the kernel, so two launches in flight - on whatever streams - never share a counter.
int* new_queue(cudaStream_t stream)
{
  int* q = nullptr;
  if (cudaMallocAsync((void**)&q, sizeof(int), stream) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  return q;
}

int ensure_stage(prx_handle* h, size_t bytes)
{
  if (bytes <= h->d_stage_bytes) {
    return PRX_OK;
  }
  if (h->d_stage != nullptr) {
    cudaFree(h->d_stage);
    h->d_stage = nullptr;
    h->d_stage_bytes = 0;
  }
  PRX_CUDA(h, cudaMalloc(&h->d_stage, bytes));
  h->d_stage_bytes = bytes;
  return PRX_OK;
}

int env_int(const char* name, int dflt)
{
  const char* s = std::getenv(name);
  return (s != nullptr && *s) ? std::atoi(s) : dflt;
}

} // namespace

// Register-resident DFMA chains: the measured FP64 roofline denominator (MEASURED_PEAKS.json has only
// HBM and bf16).  8 independent accumulators per thread, 2 flops per FMA.
__global__ void fp64_peak_kernel(double* out, const int iters, const double a, const double b)
{
  double x0 = threadIdx.x * 1e-9, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6,
         x7 = x0 + 7;
  for (int i = 0; i < iters; i++) {
    x0 = fma(x0, a, b);
    x1 = fma(x1, a, b);
    x2 = fma(x2, a, b);
    x3 = fma(x3, a, b);
    x4 = fma(x4, a, b);
    x5 = fma(x5, a, b);
    x6 = fma(x6, a, b);
    x7 = fma(x7, a, b);
  }
  const double r = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
  if (r == 123.456) {
    out[0] = r;
  }
}

extern "C" {

int prx_measure_fp64_peak(int device, double* tflops_out)
{
  if (tflops_out == nullptr) {
    return PRX_ERR_BAD_ARG;
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) {
    cudaGetLastError();
    return PRX_ERR_NO_DEVICE;
  }
  if (device < 0) {
    cudaGetDevice(&device);
  }
  DeviceGuard g(device);
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  double* d = nullptr;
  if (cudaMalloc(&d, sizeof(double)) != cudaSuccess) {
    return PRX_ERR_CUDA;
  }
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int threads = 256, blocks = sms * 8, iters = 1 << 15;
  double best = 0.0;
  for (int rep = 0; rep < 6; rep++) {
    cudaEventRecord(e0);
    fp64_peak_kernel<<<blocks, threads>>>(d, iters, 1.0000001, 1e-9);
    cudaEventRecord(e1);
    if (cudaEventSynchronize(e1) != cudaSuccess) {
      break;
    }
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    const double fl = 2.0 * 8.0 * (double)iters * threads * (double)blocks;
    const double tf = fl / (ms * 1e-3) / 1e12;
    if (rep > 0 && tf > best) {
      best = tf;
    }
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d);
  *tflops_out = best;
  return (cudaGetLastError() == cudaSuccess) ? PRX_OK : PRX_ERR_CUDA;
}

int prx_default_config(prx_config* cfg)
{
  if (cfg == nullptr) {
    return PRX_ERR_BAD_ARG;
  }
  std::memset(cfg, 0, sizeof(*cfg));
  cfg->abi_version = PRX_ABI_VERSION;
  cfg->mechanism = PRX_MECH_LIDRYER;
  cfg->reactor_kind = PRX_REACTOR_CVODE;
  cfg->reactor_type = 1;
  cfg->rtol = 1e-6;  // ReactorCvode.H:179-191 defaults
  cfg->atol = 1e-10;
  cfg->max_order = 2;
  cfg->max_nls_iters = 3;
  cfg->solve_type = PRX_SOLVE_DENSE_AJ_DIRECT;
  cfg->rk64_nsubsteps_guess = 10;
  cfg->rk64_nsubsteps_min = 5;
  cfg->rk64_nsubsteps_max = 500;
  cfg->clean_init_massfrac = 0;
  cfg->verbose = 0;
  cfg->device = -1;
  cfg->ark_method = 40;    // ReactorArkode.H: rk_method{40} = ARKODE_ZONNEVELD_5_3_4
  cfg->ark_controller = 0; // PID
  cfg->ark_use_erkstep = 1;
  return PRX_OK;
}

int prx_num_species(int mechanism)
{
  const MechOps* o = ops_for(mechanism);
  return o ? o->num_species : PRX_ERR_BAD_ARG;
}
int prx_num_reactions(int mechanism)
{
  const MechOps* o = ops_for(mechanism);
  return o ? o->num_reactions : PRX_ERR_BAD_ARG;
}
const char* prx_mechanism_name(int mechanism)
{
  const MechOps* o = ops_for(mechanism);
  return o ? o->name : nullptr;
}

int prx_create(const prx_config* cfg, prx_handle** out)
{
  if (cfg == nullptr || out == nullptr) {
    return PRX_ERR_BAD_ARG;
  }
  *out = nullptr;
  if (cfg->abi_version != PRX_ABI_VERSION) {
    return PRX_ERR_BAD_ARG;
  }
  const MechOps* ops = ops_for(cfg->mechanism);
  if (ops == nullptr) {
    return PRX_ERR_BAD_ARG;
  }
  if (cfg->reactor_kind != PRX_REACTOR_RK64 && cfg->reactor_kind != PRX_REACTOR_CVODE && cfg->reactor_kind != PRX_REACTOR_ARKODE) {
    return PRX_ERR_BAD_ARG;
  }
  ark::Table ark_tab;
  if (cfg->reactor_kind == PRX_REACTOR_ARKODE) {
    if (!(cfg->rtol > 0.0) || !(cfg->atol > 0.0)) {
      return PRX_ERR_BAD_ARG;
    }
    // ReactorArkode.cpp:25-90: an unknown ode.rk_method falls back to the default table
    if (cfg->ark_controller < 0 || cfg->ark_controller > 2) {
      return PRX_ERR_UNSUPPORTED; // the explicit-Gustafsson controller of ReactorArkode.cpp:92-116
    }
    if (!ark::table_for(cfg->ark_method, &ark_tab)) {
      ark::table_for(40, &ark_tab);
    }
  }
  if (cfg->reactor_type != 1 && cfg->reactor_type != 2) {
    return PRX_ERR_BAD_ARG; // reference: amrex::Abort("Wrong reactor type. Choose between 1 (e) or 2 (h).")
  }
  if (cfg->reactor_kind == PRX_REACTOR_CVODE) {
    if (cfg->max_order < 1 || cfg->max_order > 5 || cfg->max_nls_iters < 1 || !(cfg->rtol > 0.0) || !(cfg->atol > 0.0)) {
      return PRX_ERR_BAD_ARG;
    }
    if (cfg->solve_type != PRX_SOLVE_DENSE_AJ_DIRECT && cfg->solve_type != PRX_SOLVE_DENSE_DIRECT) {
      return PRX_ERR_UNSUPPORTED;
    }
  } else if (cfg->reactor_kind == PRX_REACTOR_RK64) {
    if (cfg->rk64_nsubsteps_guess < 1 || cfg->rk64_nsubsteps_min < 1 || cfg->rk64_nsubsteps_max < 1 || !(cfg->atol > 0.0)) {
      return PRX_ERR_BAD_ARG;
    }
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) {
    cudaGetLastError();
    return PRX_ERR_NO_DEVICE; // no CPU fallback, by design
  }
  int dev = cfg->device;
  if (dev < 0) {
    if (cudaGetDevice(&dev) != cudaSuccess) {
      return PRX_ERR_NO_DEVICE;
    }
  }
  if (dev >= ndev) {
    return PRX_ERR_BAD_ARG;
  }
  prx_handle* h = new (std::nothrow) prx_handle();
  if (h == nullptr) {
    return PRX_ERR_ALLOC;
  }
  h->cfg = *cfg;
  h->ops = ops;
  h->device = dev;
  h->ark_tab = ark_tab;
  h->ark_ctl.init(cfg->ark_controller); // ReactorArkode::init creates the SUNAdaptController once (ReactorArkode.cpp:90-116)
  h->typ_vals.assign(ops->num_species + 1, -1.0);
  h->d_stats = nullptr;
  h->n_box_streams = 0;
  h->sched.jac_defer_min_lanes = env_int("PRX_JAC_DEFER", 1);
  h->sched.lu_defer_min_lanes = env_int("PRX_LU_DEFER", 1);
  h->sched.bdf_variant = env_int("PRX_BDF_VARIANT", -1);
  h->sched.max_warps_per_sm = env_int("PRX_MAX_WARPS_PER_SM", 0);
  h->sched.coop_a_global = env_int("PRX_COOP_A_GLOBAL", -1);
  h->sched.refill_min = env_int("PRX_REFILL_MIN", -1);
  h->sched.sort_cells = env_int("PRX_SORT_CELLS", 1);
  h->sched.lu_defer_block = env_int("PRX_LU_DEFER_BLOCK", 0);
  h->sched.lu_defer_max_age = env_int("PRX_LU_DEFER_AGE", 6);
  h->sched.block_warps = env_int("PRX_BLOCK_WARPS", 0); // 0: kernel default (gen-1: both block barriers; hot/cold: single warps)
  h->last_ncells = 0;
  h->d_stage = nullptr;
  h->d_stage_bytes = 0;
  h->stage_stream = nullptr;
  h->in_stream = nullptr;
  h->out_stream = nullptr;
  h->keep_last_steps = false;
  h->err[0] = 0;
  DeviceGuard g(dev);
  {
    // the per-call scratch (cell order, Jacobian slots, cold columns) comes from the stream-ordered pool: keep the pool's
    // memory across synchronisations instead of returning it to the driver after every call
    cudaMemPool_t pool = nullptr;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess && pool != nullptr) {
      unsigned long long keep = ~0ull;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    cudaGetLastError();
  }
  if (cudaMalloc(&h->d_stats, sizeof(DevStats)) != cudaSuccess ||
      cudaMemset(h->d_stats, 0, sizeof(DevStats)) != cudaSuccess ||
      cudaStreamCreateWithFlags(&h->stage_stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaStreamCreateWithFlags(&h->in_stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaStreamCreateWithFlags(&h->out_stream, cudaStreamNonBlocking) != cudaSuccess) {
    cudaGetLastError();
    prx_destroy(h);
    return PRX_ERR_CUDA;
  }
  *out = h;
  return PRX_OK;
}

int prx_destroy(prx_handle* h)
{
  if (h == nullptr) {
    return PRX_ERR_BAD_ARG;
  }
  DeviceGuard g(h->device);
  cudaDeviceSynchronize();
  if (h->d_stats) {
    cudaFree(h->d_stats);
  }
  for (int i = 0; i < h->n_box_streams; i++) {
    cudaStreamDestroy(h->box_streams[i]);
    cudaEventDestroy(h->box_events[i]);
  }
  if (h->n_box_streams > 0) {
    cudaEventDestroy(h->box_events[8]);
  }
  if (h->d_stage) {
    cudaFree(h->d_stage);
  }
  if (h->in_stream) {
    cudaStreamDestroy(h->in_stream);
  }
  for (int i = 0; i < 3; i++) {
    if (h->host_react_streams[i]) {
      cudaStreamDestroy(h->host_react_streams[i]);
    }
  }
  if (h->out_stream) {
    cudaStreamDestroy(h->out_stream);
  }
  if (h->stage_stream) {
    cudaStreamDestroy(h->stage_stream);
  }
  delete h;
  return PRX_OK;
}

int prx_set_typ_vals(prx_handle* h, const double* typ, int n)
{
  if (h == nullptr || typ == nullptr) {
    return PRX_ERR_BAD_ARG;
  }
  if (n != h->ops->num_species + 1) {
    return fail(h, PRX_ERR_BAD_ARG, "prx_set_typ_vals: n must be NUM_SPECIES+1");
  }
  for (int k = 0; k < n; k++) {
    h->typ_vals[k] = typ[k];
  }
  return PRX_OK;
}

const char* prx_last_error(const prx_handle* h) { return h ? h->err : "null handle"; }

// ---- cell ordering -----------------------------------------------------------------------------
// The kernels draw cells from a work queue; handing them out SORTED BY TEMPERATURE puts thermochemically
// similar cells (similar step counts, Newton iterations and Jacobian refreshes) into the same warp, which is
// what keeps one-cell-per-thread integrators converged on flame-like fields where neighbours in index space
// are unrelated (bench.py --states pmf: 319 -> see profiles/README.md).  Each cell's arithmetic does not depend
// on the order.  Key = T as float; stable radix sort (cub), so equal temperatures keep index order.
__global__ void order_keys_kernel(const BoxArgs a, const int neq, const float sign, float* __restrict__ keys, int* __restrict__ idx)
{
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= a.ncells) {
    return;
  }
  double T;
  if (a.layout == LAYOUT_ARRAY4) {
    const int i = c % a.len[0];
    const int r = c / a.len[0];
    const int j = r % a.len[1];
    const int k = r / a.len[1];
    T = a.T.p[(long long)i + (long long)j * a.T.js + (long long)k * a.T.ks];
  } else if (a.layout == LAYOUT_STATE) {
    const int i = c % a.len[0];
    const int r = c / a.len[0];
    const int j = r % a.len[1];
    const int k = r / a.len[1];
    T = a.sold.p[(long long)i + (long long)j * a.sold.js + (long long)k * a.sold.ks + (long long)a.ix.utemp * a.sold.ns];
  } else {
    T = a.y[(long long)c * neq + (neq - 1)];
  }
  keys[c] = sign * (float)T; // sign = -1: hottest cells first (PRX_SORT_CELLS=2)
  idx[c] = c;
}

// returns a stream-ordered allocation holding [order | scratch]; the caller frees it (cudaFreeAsync) after the launch
static cudaError_t build_order(const BoxArgs& a, int neq, bool descending, cudaStream_t stream, void** block_out, const int** order_out)
{
  const size_t n = (size_t)a.ncells;
  size_t temp_bytes = 0;
  cudaError_t e = cub::DeviceRadixSort::SortPairs(
    nullptr, temp_bytes, (const float*)nullptr, (float*)nullptr, (const int*)nullptr, (int*)nullptr, (int)n, 0, 32, stream);
  if (e != cudaSuccess) {
    return e;
  }
  const size_t off_keys_in = n * sizeof(int), off_keys_out = off_keys_in + n * sizeof(float),
               off_idx_in = off_keys_out + n * sizeof(float), off_temp = (off_idx_in + n * sizeof(int) + 255) & ~(size_t)255;
  char* block = nullptr;
  e = cudaMallocAsync((void**)&block, off_temp + temp_bytes, stream);
  if (e != cudaSuccess) {
    return e;
  }
  int* order = (int*)block;
  float* keys_in = (float*)(block + off_keys_in);
  float* keys_out = (float*)(block + off_keys_out);
  int* idx_in = (int*)(block + off_idx_in);
  order_keys_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(a, neq, descending ? -1.0f : 1.0f, keys_in, idx_in);
  e = cudaGetLastError();
  if (e == cudaSuccess) {
    e = cub::DeviceRadixSort::SortPairs(block + off_temp, temp_bytes, keys_in, keys_out, idx_in, order, (int)n, 0, 32, stream);
  }
  if (e != cudaSuccess) {
    cudaFreeAsync(block, stream);
    return e;
  }
  *block_out = block;
  *order_out = order;
  return cudaSuccess;
}

// one NVTX range per react call (the reference wraps react in BL_PROFILE("Pele::ReactorCvode::react()"))
struct NvtxRange
{
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
};

static int react_args(prx_handle* h, const BoxArgs& a, cudaStream_t stream)
{
  NvtxRange nvtx(h->cfg.reactor_kind == PRX_REACTOR_RK64 ? "prx::ReactorRK64::react" : "prx::ReactorCvode::react");
  DeviceGuard g(h->device);
  // per-launch step sum (for the value ReactorBase::react returns)
  if (!h->keep_last_steps) {
    PRX_CUDA(h, cudaMemsetAsync(&h->d_stats->v[ST_LAST_STEPS], 0, 2 * sizeof(unsigned long long), stream)); // + ST_LAST_MAXRHS
    h->last_ncells = 0;
  }
  h->last_layout = a.layout;
  h->last_ncells += a.ncells;
  cudaError_t e;
  BoxArgs b = a;
  void* order_block = nullptr;
  b.order = nullptr;
  if (h->sched.sort_cells && a.ncells >= 4096) {
    e = build_order(a, h->ops->num_species + 1, h->sched.sort_cells == 2, stream, &order_block, &b.order);
    if (e != cudaSuccess) {
      return fail(h, PRX_ERR_CUDA, "cell ordering", e);
    }
  }
  SchedParams sp = h->sched;
  if (sp.refill_min < 0) { // default: batch the refills of a warp only when the cells arrive unsorted
    sp.refill_min = (b.order != nullptr) ? 1 : 16;
  }
  int* queue = new_queue(stream);
  if (queue == nullptr) {
    if (order_block != nullptr) {
      cudaFreeAsync(order_block, stream);
    }
    return fail(h, PRX_ERR_ALLOC, "work-queue counter");
  }
  if (h->cfg.reactor_kind == PRX_REACTOR_RK64) {
    e = h->ops->launch_rk64(b, h->cfg, queue, h->d_stats, stream);
  } else {
    e = h->ops->launch_bdf(b, h->cfg, h->typ_vals.data(), queue, h->d_stats, sp, stream);
  }
  cudaFreeAsync(queue, stream);
  if (order_block != nullptr) {
    cudaFreeAsync(order_block, stream); // stream-ordered: released after the integrator kernel
  }
  if (e != cudaSuccess) {
    return fail(h, PRX_ERR_CUDA, "kernel launch", e);
  }
  return PRX_OK;
}

// ---- ReactorArkode's 1-D react (ReactorArkode.cpp:266-380): the caller's cell-major vectors go into the integrator's CYOrder
// columns as they are (no initial temperature solve: T is component N of each cell), and come back the same way; the energy
// is advanced by its external source.
__global__ void ark_flat_in_kernel(
  const int nc, const int N, const double* __restrict__ y, const double* __restrict__ ysrc, const double* __restrict__ e,
  const double* __restrict__ esrc, double* __restrict__ yn, double* __restrict__ rYsrc, double* __restrict__ rhoe,
  double* __restrict__ rhoesrc)
{
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= nc) {
    return;
  }
  for (int k = 0; k <= N; k++) {
    yn[(long long)k * nc + c] = y[(long long)c * (N + 1) + k];
  }
  for (int k = 0; k < N; k++) {
    rYsrc[(long long)k * nc + c] = ysrc[(long long)c * N + k];
  }
  rhoe[c] = e[c];
  rhoesrc[c] = esrc[c];
}

__global__ void ark_flat_out_kernel(
  const int nc, const int N, const double dt, const double* __restrict__ yn, const double* __restrict__ esrc,
  double* __restrict__ y, double* __restrict__ e, double* __restrict__ fc, const double nfe)
{
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= nc) {
    return;
  }
  for (int k = 0; k <= N; k++) {
    y[(long long)c * (N + 1) + k] = yn[(long long)k * nc + c];
  }
  e[c] = e[c] + dt * esrc[c];
  if (fc != nullptr) {
    fc[c] = nfe;
  }
}

// ---- ReactorArkode::react, Array4 (PP/Source/Reactions/ReactorArkode.cpp:135-260): see prx_arkode.cuh ---------------------
static int react_arkode(prx_handle* h, const BoxArgs& a, cudaStream_t stream)
{
  NvtxRange nvtx("prx::ReactorArkode::react");
  DeviceGuard g(h->device);
  const MechOps* ops = h->ops;
  const ark::Table& B = h->ark_tab;
  const int N = ops->num_species, neq = N + 1;
  const long long nc = a.ncells, n = nc * neq;
  const double uround = 2.220446049250313e-16;
  // device vectors: yn, ycur, yold, fold, ewt, tmp, F[stages] (neq columns each), rYsrc (N), rhoe, rhoesrc, atol (neq), scratch
  const int nvec = 6 + B.stages;
  const size_t doubles = (size_t)nvec * n + (size_t)N * nc + 2 * (size_t)nc + (size_t)neq + 2 * ark::RG + 2;
  double* buf = nullptr;
  if (cudaMallocAsync((void**)&buf, doubles * sizeof(double) + (size_t)nc * sizeof(long long), stream) != cudaSuccess) {
    cudaGetLastError();
    return fail(h, PRX_ERR_ALLOC, "ReactorArkode work vectors");
  }
  double* yn = buf;
  double* ycur = yn + n;
  double* yold = ycur + n;
  double* fold = yold + n;
  double* ewt = fold + n;
  double* tmp = ewt + n;
  double* F[13] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  for (int sidx = 0; sidx < B.stages; sidx++) {
    F[sidx] = tmp + (size_t)(sidx + 1) * n;
  }
  double* rYsrc = tmp + (size_t)(B.stages + 1) * n;
  double* rhoe = rYsrc + (size_t)N * nc;
  double* rhoesrc = rhoe + nc;
  double* atolv = rhoesrc + nc;
  double* part = atolv + neq;
  double* red = part + 2 * ark::RG;
  long long* fcunt = reinterpret_cast<long long*>(buf + doubles);
  int rc = PRX_OK;
  long long nfe = 0;
  double host_red[2] = {0.0, 0.0};
  auto fetch = [&]() { // reduce the block partials and bring (sum, max) to the host
    ark::finish_kernel<<<1, 32, 0, stream>>>(part, red);
    cudaMemcpyAsync(host_red, red, 2 * sizeof(double), cudaMemcpyDeviceToHost, stream);
    return cudaStreamSynchronize(stream);
  };
  auto wrms = [&](double sumsq) { return std::sqrt(sumsq / (double)n); };
  auto rhs_at = [&](double t, const double* y, double* f) { // plain f(t, y)
    ArkStageArgs s;
    std::memset(&s, 0, sizeof(s));
    s.ncells = a.ncells;
    s.reactor_type = h->cfg.reactor_type;
    s.nterms = 0;
    s.t = t;
    s.yn = y;
    s.ycur = nullptr;
    s.fout = f;
    s.rhoe = rhoe;
    s.rhoesrc = rhoesrc;
    s.rYsrc = rYsrc;
    nfe++;
    return ops->ark_stage(s, stream);
  };
#define ARK_CUDA(call)                                         \
  do {                                                         \
    const cudaError_t e_ = (call);                             \
    if (e_ != cudaSuccess && rc == PRX_OK) {                   \
      rc = fail(h, PRX_ERR_CUDA, #call, e_);                   \
    }                                                          \
  } while (0)
  do {
    // absolute tolerances as set_sundials_solver_tols builds them (ReactorUtils.H:119-141)
    std::vector<double> atol_h(neq);
    for (int k = 0; k < neq; k++) {
      atol_h[k] = (h->typ_vals[0] > 0.0) ? h->typ_vals[k] * h->cfg.atol : h->cfg.atol;
    }
    ARK_CUDA(cudaMemcpyAsync(atolv, atol_h.data(), neq * sizeof(double), cudaMemcpyHostToDevice, stream));
    // flatten (CYOrder), incl. the initial T solve of box_flatten
    if (a.layout == LAYOUT_FLAT) {
      ark_flat_in_kernel<<<(unsigned)((nc + 255) / 256), 256, 0, stream>>>((int)nc, N, a.y, a.ysrc, a.e, a.esrc, yn, rYsrc, rhoe, rhoesrc);
      ARK_CUDA(cudaGetLastError());
    } else {
      ARK_CUDA(ops->flatten(a, h->cfg, 1, yn, rYsrc, rhoe, rhoesrc, stream));
    }
    const double t0 = a.time, tout = a.time + a.dt;
    double tn = t0, tcur = t0;
    // ---- arkInitialSetup: error weights, f(t0, y0), initial step
    ark::ewt_kernel<<<ark::RG, ark::RB, 0, stream>>>(n, nc, yn, h->cfg.rtol, atolv, ewt, part);
    ARK_CUDA(fetch());
    if (rc != PRX_OK) {
      break;
    }
    if (host_red[1] != 0.0) {
      rc = fail(h, PRX_ERR_BAD_ARG, "ReactorArkode: a component of the initial error weight is not positive (ARK_ILL_INPUT)");
      break;
    }
    double yn_sumsq = host_red[0]; // sum (yn ewt)^2 of the current error weights: the "too much accuracy" test of every step
    const double tdiff = tout - tcur;
    const double tdist = std::fabs(tdiff), tround = uround * std::fmax(std::fabs(tcur), std::fabs(tout));
    if (tdiff == 0.0 || tdist < 2.0 * tround) {
      rc = fail(h, PRX_ERR_BAD_ARG, "ReactorArkode: tout too close to t0 (ARK_TOO_CLOSE)");
      break;
    }
    const double sign = (tdiff > 0.0) ? 1.0 : -1.0;
    ARK_CUDA(rhs_at(tn, yn, F[0])); // step_fullrhs(ARK_FULLRHS_START): fn = F[0]
    double hstep;
    {
      // arkHin
      const double hlb = 100.0 * tround;
      ark::hub_kernel<<<ark::RG, ark::RB, 0, stream>>>(n, yn, F[0], ewt, part);
      ARK_CUDA(fetch());
      double hub = 0.1 * tdist;
      if (hub * host_red[1] > 1.0) {
        hub = 1.0 / host_red[1];
      }
      double hg = std::sqrt(hlb * hub);
      if (hub < hlb) {
        hstep = sign * hg;
      } else {
        double hnew = hg;
        for (int count1 = 1; count1 <= 4; count1++) {
          const double hgs = hg * sign;
          ark::axpy_kernel<<<ark::RG, ark::RB, 0, stream>>>(n, hgs, F[0], yn, ycur);
          ARK_CUDA(rhs_at(tcur + hgs, ycur, tmp));
          ark::ydd_kernel<<<ark::RG, ark::RB, 0, stream>>>(n, 1.0 / hgs, tmp, F[0], ewt, part);
          ARK_CUDA(fetch());
          const double yddnrm = wrms(host_red[0]);
          hnew = (yddnrm * hub * hub > 2.0) ? std::sqrt(2.0 / yddnrm) : std::sqrt(hg * hub);
          if (count1 == 4) {
            break;
          }
          const double hrat = hnew / hg;
          if (hrat > 0.5 && hrat < 2.0) {
            break;
          }
          if (count1 > 1 && hrat > 2.0) {
            hnew = hg;
            break;
          }
          hg = hnew;
        }
        double h0 = 0.5 * hnew;
        h0 = h0 < hlb ? hlb : h0;
        h0 = h0 > hub ? hub : h0;
        hstep = sign * h0;
      }
    }
    if (rc != PRX_OK) {
      break;
    }
    // ERKStep overrides ARKODE's adaptivity defaults (erkStep_SetDefaults, arkode_erkstep_io.c:310-312: safety 0.99, growth 25);
    // ARKStep with an explicit table (ode.use_erkstep = 0) runs with ARKODE's own (arkode_adapt_impl.h:41-42: 0.96, 20) -
    // measured against the reference: the two differ in nothing else
    const double ark_safety = h->cfg.ark_use_erkstep ? 0.99 : 0.96, ark_growth = h->cfg.ark_use_erkstep ? 25.0 : 20.0;
    double etamax = 10000.0; // etamx1: first step
    double eta = 1.0, hprime = hstep;
    bool initsetup = true, fn_is_current = true;
    bool stiffly_accurate = true; // ARKodeButcherTable_IsStifflyAccurate: b == last row of A
    for (int j = 0; j < B.stages; j++) {
      stiffly_accurate = stiffly_accurate && (std::fabs(B.b[j] - B.A[B.stages - 1][j]) <= 100.0 * uround);
    }
    double told_i = tcur, tnew_i = tcur, h_i = 0.0; // Hermite interpolation data (yold, fold in device vectors)
    long nstloc = 0;
    bool done = false;
    while (!done && rc == PRX_OK) {
      if (!initsetup) {
        ark::ewt_kernel<<<ark::RG, ark::RB, 0, stream>>>(n, nc, yn, h->cfg.rtol, atolv, ewt, part);
        ARK_CUDA(fetch());
        if (host_red[1] != 0.0) {
          break; // ARK_ILL_INPUT: the reference returns yn (ReactorArkode ignores the flag)
        }
        yn_sumsq = host_red[0];
      }
      if (nstloc >= 500) { // MXSTEP_DEFAULT: ARK_TOO_MUCH_WORK, the reference carries on with yn at tcur
        h->ark_too_much_work++;
        break;
      }
      if (uround * wrms(yn_sumsq) > 1.0) {
        break; // ARK_TOO_MUCH_ACC
      }
      if (hprime != hstep) {
        hstep = hstep * eta;
      }
      int nef = 0;
      double dsm = 0.0;
      bool step_ok = false;
      for (;;) { // attempts
        if (!fn_is_current) { // erkStep_FullRHS(ARK_FULLRHS_END)
          if (stiffly_accurate) {
            ARK_CUDA(cudaMemcpyAsync(F[0], F[B.stages - 1], n * sizeof(double), cudaMemcpyDeviceToDevice, stream));
          } else {
            ARK_CUDA(rhs_at(tn, yn, F[0]));
          }
          fn_is_current = true;
        }
        for (int is = 1; is < B.stages; is++) {
          ArkStageArgs s;
          std::memset(&s, 0, sizeof(s));
          s.ncells = a.ncells;
          s.reactor_type = h->cfg.reactor_type;
          s.t = tn + B.c[is] * hstep;
          for (int js = 0; js < is; js++) {
            if (B.A[is][js] != 0.0) { // zero terms add an exact zero in the reference's linear combination
              s.coef[s.nterms] = hstep * B.A[is][js];
              s.F[s.nterms] = F[js];
              s.nterms++;
            }
          }
          s.yn = yn;
          s.ycur = nullptr; // N_VLinearCombination writes ark_mem->ycur; nothing reads it before the next stage overwrites it
          s.fout = F[is];
          s.rhoe = rhoe;
          s.rhoesrc = rhoesrc;
          s.rYsrc = rYsrc;
          if (s.nterms == 0) { // cannot happen for a consistent table with c > 0; keep the kernel's contract
            s.coef[0] = 0.0;
            s.F[0] = F[0];
            s.nterms = 1;
          }
          nfe++;
          ARK_CUDA(ops->ark_stage(s, stream));
        }
        ark::CombineArgs ca;
        std::memset(&ca, 0, sizeof(ca));
        ca.stages = B.stages;
        for (int j = 0; j < B.stages; j++) {
          ca.cb[j] = hstep * B.b[j];
          ca.ce[j] = hstep * (B.b[j] - B.d[j]);
          ca.F[j] = F[j];
        }
        ark::combine_kernel<<<ark::RG, ark::RB, 0, stream>>>(n, ca, yn, ewt, ycur, part);
        ARK_CUDA(fetch());
        if (rc != PRX_OK) {
          break;
        }
        dsm = wrms(host_red[0]);
        // arkCheckTemporalError -> arkAdapt
        {
          const int controller_order = B.p + (-1); // pq = 0: embedding order + ADJUST
          double h_acc = h->ark_ctl.estimate(hstep, controller_order, dsm);
          const double int_dir = hstep / std::fabs(hstep);
          double h_cfl = 1.0e30 * std::fabs(hstep); // arkExpStab returns 0: no explicit stability limit
          h_acc *= ark_safety;
          h_cfl *= 0.5 * int_dir;
          h_acc = int_dir * std::fmin(std::fabs(h_acc), std::fabs(etamax * hstep));
          h_acc = int_dir * std::fmax(std::fabs(h_acc), std::fabs(0.1 * hstep));
          h_acc = int_dir * std::fmin(std::fabs(h_acc), std::fabs(h_cfl));
          if (dsm <= 1.0) {
            if ((std::fabs(h_acc) > std::fabs(hstep * 1.0 * 0.999999)) && (std::fabs(h_acc) < std::fabs(hstep * 1.5 * 1.000001))) {
              h_acc = hstep;
            }
          }
          eta = h_acc / hstep;
          eta = std::fmax(eta, 0.0 / std::fabs(hstep));        // hmin = 0
          eta /= std::fmax(1.0, std::fabs(hstep) * 0.0 * eta); // hmax_inv = 0
        }
        eta = std::fmin(eta, etamax);
        eta = std::fmax(eta, 0.0);
        eta /= std::fmax(1.0, 0.0);
        if (dsm <= 1.0) {
          step_ok = true;
          break;
        }
        nef++;
        h->ark_netf++;
        if (nef == 7) { // MAXNEF: ARK_ERR_FAILURE
          break;
        }
        etamax = 1.0;
        if (nef >= 2) {
          eta = std::fmin(eta, 0.3); // ETAMXF from the SMALL_NEF-th failure on
        }
        eta = std::fmin(eta, etamax);
        hstep *= eta;
        hprime = hstep;
      }
      if (!step_ok || rc != PRX_OK) {
        break; // the reference's arkHandleFailure: yout = yn
      }
      // arkCompleteStep
      tcur = tn + hstep;
      ARK_CUDA(cudaMemcpyAsync(yold, yn, n * sizeof(double), cudaMemcpyDeviceToDevice, stream)); // arkInterpUpdate_Hermite
      ARK_CUDA(cudaMemcpyAsync(fold, F[0], n * sizeof(double), cudaMemcpyDeviceToDevice, stream));
      told_i = tnew_i;
      tnew_i = tcur;
      h_i = hstep;
      std::swap(yn, ycur); // yn <- ycur
      fn_is_current = false;
      h->ark_ctl.update(hstep, dsm);
      h->ark_nst++;
      tn = tcur;
      hprime = hstep * eta;
      etamax = ark_growth;
      initsetup = false;
      nstloc++;
      if ((tcur - tout) * hstep >= 0.0) { // ARK_NORMAL: tout reached, interpolate back (ARKodeGetDky)
        const double tau = (tout - tcur) / hstep;
        if (stiffly_accurate) {
          ARK_CUDA(cudaMemcpyAsync(F[0], F[B.stages - 1], n * sizeof(double), cudaMemcpyDeviceToDevice, stream));
        } else {
          ARK_CUDA(rhs_at(tn, yn, F[0])); // fn at the new point
        }
        fn_is_current = true;
        const double tau2 = tau * tau, tau3 = tau * tau2;
        const int deg = B.q - 1 < 5 ? B.q - 1 : 5; // interp degree min(5, q - 1)
        if (deg >= 5) {
          // quintic (arkode_interp.c:648-690).  The reference evaluates the quartic interpolant at tau = -1/3 and the right-hand
          // side there, then the quartic at tau = -2/3 and the right-hand side there (fb); each quartic evaluation first
          // evaluates the cubic at tau = -1/3 and stores ITS right-hand side in fa - so the fa of the final formula is f at
          // the cubic's point, and f at the quartic's -1/3 point is computed, counted and overwritten.  Four right-hand sides
          // are counted; the two whose results the reference discards are not launched.
          const double ta = -1.0 / 3.0, ta2 = ta * ta, ta3 = ta * ta2;
          ark::interp_kernel<<<ark::RG, ark::RB, 0, stream>>>(
            n, 3.0 * ta2 + 2.0 * ta3, 1.0 - 3.0 * ta2 - 2.0 * ta3, h_i * (ta2 + ta3), h_i * (ta + 2.0 * ta2 + ta3), yold, yn, fold,
            F[0], ycur);
          ARK_CUDA(rhs_at(tnew_i - h_i / 3.0, ycur, tmp)); // fa
          nfe += 2;                                         // the discarded pair (cubic point again, quartic(-1/3) point)
          const double tb = -2.0 / 3.0, tb2 = tb * tb, tb3 = tb * tb2, tb4 = tb * tb3;
          ark::interp5_kernel<<<ark::RG, ark::RB, 0, stream>>>(
            n, -6.0 * tb2 - 16.0 * tb3 - 9.0 * tb4, 1.0 + 6.0 * tb2 + 16.0 * tb3 + 9.0 * tb4,
            h_i * 0.25 * (-5.0 * tb2 - 14.0 * tb3 - 9.0 * tb4), h_i * (tb + 2.0 * tb2 + tb3),
            h_i * 27.0 * 0.25 * (-tb4 - 2.0 * tb3 - tb2), yold, yn, fold, F[0], tmp, ycur);
          ARK_CUDA(rhs_at(tnew_i - h_i * 2.0 / 3.0, ycur, F[1])); // fb (the stage vectors are free after the step)
          const double tau4 = tau * tau3, tau5 = tau * tau4;
          const double a0 = 54.0 * tau5 + 135.0 * tau4 + 110.0 * tau3 + 30.0 * tau2;
          ark::interp6_kernel<<<ark::RG, ark::RB, 0, stream>>>(
            n, a0, 1.0 - a0, h_i / 4.0 * (27.0 * tau5 + 63.0 * tau4 + 49.0 * tau3 + 13.0 * tau2),
            h_i / 4.0 * (27.0 * tau5 + 72.0 * tau4 + 67.0 * tau3 + 26.0 * tau2 + 4.0 * tau),
            h_i / 4.0 * (81.0 * tau5 + 189.0 * tau4 + 135.0 * tau3 + 27.0 * tau2),
            h_i / 4.0 * (81.0 * tau5 + 216.0 * tau4 + 189.0 * tau3 + 54.0 * tau2), yold, yn, fold, F[0], tmp, F[1], ycur);
        } else if (deg >= 4) {
          // quartic (arkode_interp.c:578-600): the cubic interpolant at tau = -1/3, one more right-hand side there (counted in
          // nfe like the reference's step_fullrhs), then five terms
          const double ta = -1.0 / 3.0, ta2 = ta * ta, ta3 = ta * ta2;
          ark::interp_kernel<<<ark::RG, ark::RB, 0, stream>>>(
            n, 3.0 * ta2 + 2.0 * ta3, 1.0 - 3.0 * ta2 - 2.0 * ta3, h_i * (ta2 + ta3), h_i * (ta + 2.0 * ta2 + ta3), yold, yn, fold,
            F[0], ycur);
          ARK_CUDA(rhs_at(tnew_i - h_i / 3.0, ycur, tmp));
          const double tau4 = tau * tau3;
          ark::interp5_kernel<<<ark::RG, ark::RB, 0, stream>>>(
            n, -6.0 * tau2 - 16.0 * tau3 - 9.0 * tau4, 1.0 + 6.0 * tau2 + 16.0 * tau3 + 9.0 * tau4,
            h_i * 0.25 * (-5.0 * tau2 - 14.0 * tau3 - 9.0 * tau4), h_i * (tau + 2.0 * tau2 + tau3),
            h_i * 27.0 * 0.25 * (-tau4 - 2.0 * tau3 - tau2), yold, yn, fold, F[0], tmp, ycur);
        } else if (deg >= 3) {
          ark::interp_kernel<<<ark::RG, ark::RB, 0, stream>>>(
            n, 3.0 * tau2 + 2.0 * tau3, 1.0 - 3.0 * tau2 - 2.0 * tau3, h_i * (tau2 + tau3), h_i * (tau + 2.0 * tau2 + tau3), yold,
            yn, fold, F[0], ycur);
        } else if (deg == 2) {
          ark::interp_kernel<<<ark::RG, ark::RB, 0, stream>>>(n, tau2, 1.0 - tau2, h_i * (tau2 + tau), 0.0, yold, yn, F[0], F[0], ycur);
        } else {
          ark::interp_kernel<<<ark::RG, ark::RB, 0, stream>>>(n, -tau, 1.0 + tau, 0.0, 0.0, yold, yn, F[0], F[0], ycur);
        }
        std::swap(yn, ycur); // the result vector
        done = true;
      }
    }
    (void)told_i;
    if (rc != PRX_OK) {
      break;
    }
    // unflatten with FC = nfe in every cell (ReactorArkode.cpp:237-241); react returns nfe
    if (a.layout == LAYOUT_FLAT) {
      ark_flat_out_kernel<<<(unsigned)((nc + 255) / 256), 256, 0, stream>>>((int)nc, N, a.dt, yn, a.esrc, a.y, a.e, a.fc, (double)nfe);
      ARK_CUDA(cudaGetLastError());
    } else {
      ark::fill_ll_kernel<<<ark::RG, ark::RB, 0, stream>>>(nc, nfe, fcunt);
      BoxArgs b = a;
      ARK_CUDA(ops->unflatten(b, h->cfg, 1, yn, rhoe, fcunt, stream));
    }
    h->ark_last_nfe = nfe;
    h->ark_nfe += nfe;
    h->ark_cells += nc;
  } while (0);
#undef ARK_CUDA
  cudaFreeAsync(buf, stream);
  return rc;
}

int prx_react_box(
  prx_handle* h, const int lo[3], const int hi[3], const prx_array4* rY, const prx_array4* rYsrc,
  const prx_array4* T, const prx_array4* rE, const prx_array4* rEsrc, const prx_array4* FC,
  const prx_array4i* mask, double dt, double time, void* cuda_stream)
{
  if (h == nullptr || lo == nullptr || hi == nullptr) {
    return PRX_ERR_BAD_ARG;
  }
  const int N = h->ops->num_species;
  BoxArgs a;
  std::memset(&a, 0, sizeof(a));
  a.layout = LAYOUT_ARRAY4;
  long long nc = 1;
  for (int d = 0; d < 3; d++) {
    a.len[d] = hi[d] - lo[d] + 1;
    if (a.len[d] <= 0) {
      a.len[d] = 0;
    }
    nc *= a.len[d];
  }
  if (nc == 0) {
    return PRX_OK; // empty box: nothing to do (amrex::ParallelFor over an empty box)
  }
  if (nc > 2147483647LL) {
    return fail(h, PRX_ERR_BAD_ARG, "box too large (int32 cell index)");
  }
  a.ncells = (int)nc;
  if (!view_of(rY, lo, N, &a.rY, hi) || !view_of(rYsrc, lo, N, &a.rYsrc, hi) || !view_of(T, lo, 1, &a.T, hi) ||
      !view_of(rE, lo, 1, &a.rE, hi) || !view_of(rEsrc, lo, 1, &a.rEsrc, hi) || !view_of(FC, lo, 1, &a.FC, hi)) {
    return fail(h, PRX_ERR_BAD_ARG, "prx_react_box: null/short Array4 or box outside the array");
  }
  if (mask != nullptr && mask->p != nullptr) {
    for (int d = 0; d < 3; d++) {
      if (lo[d] < mask->begin[d] || hi[d] >= mask->end[d] || mask->ncomp < 1) {
        return fail(h, PRX_ERR_BAD_ARG, "prx_react_box: box outside the mask array");
      }
    }
    a.mask.js = mask->jstride;
    a.mask.ks = mask->kstride;
    a.mask.ns = mask->nstride;
    a.mask.p = mask->p + (lo[0] - mask->begin[0]) + (long long)(lo[1] - mask->begin[1]) * mask->jstride +
               (long long)(lo[2] - mask->begin[2]) * mask->kstride;
  }
  a.dt = dt;
  a.time = time;
  if (h->cfg.reactor_kind == PRX_REACTOR_ARKODE) {
    return react_arkode(h, a, (cudaStream_t)cuda_stream); // the mask is ignored, as in the reference (ReactorArkode.cpp:143)
  }
  return react_args(h, a, (cudaStream_t)cuda_stream);
}

// ---- fused PeleC::react_state body (PC/Source/React.cpp:181-343) -----------------------------------------------
static int react_state_one(
  prx_handle* h, const int lo[3], const int hi[3], const prx_array4* sold, const prx_array4* snew, const prx_array4* nonrs,
  const prx_array4* I_R, const prx_array4* work, const prx_state_index* ix, double dt, int do_update, cudaStream_t stream)
{
  const int N = h->ops->num_species;
  if (ix->nvar < ix->ufs + N || ix->urho < 0 || ix->ufs < 0) {
    return fail(h, PRX_ERR_BAD_ARG, "prx_react_state_box: state index table does not fit NUM_SPECIES");
  }
  BoxArgs a;
  std::memset(&a, 0, sizeof(a));
  a.layout = LAYOUT_STATE;
  long long nc = 1;
  for (int d = 0; d < 3; d++) {
    a.len[d] = hi[d] - lo[d] + 1;
    if (a.len[d] <= 0) {
      return PRX_OK; // empty box
    }
    nc *= a.len[d];
  }
  if (nc > 2147483647LL) {
    return fail(h, PRX_ERR_BAD_ARG, "box too large (int32 cell index)");
  }
  a.ncells = (int)nc;
  if (!view_of(sold, lo, ix->nvar, &a.sold, hi) || !view_of(snew, lo, ix->nvar, &a.snew, hi) ||
      !view_of(nonrs, lo, ix->nvar, &a.nonrs, hi) || !view_of(I_R, lo, N + 2, &a.IR, hi)) {
    return fail(h, PRX_ERR_BAD_ARG, "prx_react_state_box: null/short Array4 or box outside the array");
  }
  if (work != nullptr && work->p != nullptr && !view_of(work, lo, 1, &a.FC, hi)) {
    return fail(h, PRX_ERR_BAD_ARG, "prx_react_state_box: box outside the work array");
  }
  a.ix.urho = ix->urho;
  a.ix.umx = ix->umx;
  a.ix.umy = ix->umy;
  a.ix.umz = ix->umz;
  a.ix.ueden = ix->ueden;
  a.ix.ueint = ix->ueint;
  a.ix.utemp = ix->utemp;
  a.ix.ufs = ix->ufs;
  a.do_update = do_update;
  a.rf_on = h->rf_on;
  if (h->rf_on) {
    a.rf_axis = h->rf_axis;
    a.rf_omega = h->rf_omega;
    for (int d = 0; d < 3; d++) {
      a.lo[d] = lo[d];
      a.rf_axloc[d] = h->rf_axloc[d];
      a.rf_problo[d] = h->rf_problo[d];
      a.rf_dx[d] = h->rf_dx[d];
    }
  }
  a.dt = dt;
  a.time = 0.0; // React.cpp:171: current_time = 0
  const bool fused = (h->cfg.reactor_kind == PRX_REACTOR_RK64) ||
                     (h->cfg.reactor_kind == PRX_REACTOR_CVODE && h->ops->bdf_is_fused(h->sched));
  if (fused) {
    return react_args(h, a, stream);
  }
  // unfused form (cooperative kernel): pack -> Array4 react on stream-ordered scratch -> unpack; still one call, no
  // MultiFab temporaries on the caller's side.  Scratch = [rY N | T | rE | rEsrc | FC] x ncells doubles.
  DeviceGuard g(h->device);
  double* scratch = nullptr;
  cudaError_t e = cudaMallocAsync((void**)&scratch, sizeof(double) * (size_t)(N + 4) * (size_t)nc, stream);
  if (e != cudaSuccess) {
    return fail(h, PRX_ERR_ALLOC, "react_state scratch", e);
  }
  auto sv = [&](double* p) {
    ArrView v;
    v.p = p;
    v.js = a.len[0];
    v.ks = (long long)a.len[0] * a.len[1];
    v.ns = nc;
    return v;
  };
  a.rY = sv(scratch);
  a.T = sv(scratch + (size_t)N * nc);
  a.rE = sv(scratch + (size_t)(N + 1) * nc);
  a.rEsrc = sv(scratch + (size_t)(N + 2) * nc);
  if (a.FC.p == nullptr) {
    a.FC = sv(scratch + (size_t)(N + 3) * nc);
  }
  a.rYsrc = a.nonrs; // the species rows of the non-reacting source, in place (read only)
  a.rYsrc.p += (long long)ix->ufs * a.nonrs.ns;
  int rc = PRX_OK;
  e = h->ops->state_pack(a, stream);
  if (e == cudaSuccess) {
    BoxArgs b = a;
    b.layout = LAYOUT_ARRAY4;
    rc = (h->cfg.reactor_kind == PRX_REACTOR_ARKODE) ? react_arkode(h, b, stream) : react_args(h, b, stream);
    if (rc == PRX_OK) {
      e = h->ops->state_unpack(a, h->cfg, stream);
    }
  }
  cudaFreeAsync(scratch, stream);
  if (rc != PRX_OK) {
    return rc;
  }
  return e == cudaSuccess ? PRX_OK : fail(h, PRX_ERR_CUDA, "react_state pack/unpack", e);
}

int prx_set_rotating_frame(
  prx_handle* h, int on, double omega, int axis, const double axis_loc[3], const double prob_lo[3], const double dx[3])
{
  if (h == nullptr || (on && (axis < 0 || axis > 2 || !axis_loc || !prob_lo || !dx))) {
    return PRX_ERR_BAD_ARG;
  }
  h->rf_on = on ? 1 : 0;
  if (on) {
    h->rf_omega = omega;
    h->rf_axis = axis;
    for (int d = 0; d < 3; d++) {
      h->rf_axloc[d] = axis_loc[d];
      h->rf_problo[d] = prob_lo[d];
      h->rf_dx[d] = dx[d];
    }
  }
  return PRX_OK;
}

int prx_react_state_box(
  prx_handle* h, const int lo[3], const int hi[3], const prx_array4* sold, const prx_array4* snew, const prx_array4* nonrs,
  const prx_array4* I_R, const prx_array4* work, const prx_state_index* ix, double dt, int do_update, void* cuda_stream)
{
  if (h == nullptr || lo == nullptr || hi == nullptr || ix == nullptr) {
    return PRX_ERR_BAD_ARG;
  }
  return react_state_one(h, lo, hi, sold, snew, nonrs, I_R, work, ix, dt, do_update, (cudaStream_t)cuda_stream);
}

// ---- PeleC::computeTemp's kernel (PC/Source/PeleC.cpp:2254-2272 -> pc_cmpTemp, Utilities.H:70-84) -----------------------
int prx_compute_temp_boxes(
  prx_handle* h, int nboxes, const int* lo, const int* hi, const prx_array4* state, const prx_state_index* ix, void* cuda_stream)
{
  if (h == nullptr || nboxes < 0 || ix == nullptr || (nboxes > 0 && (!lo || !hi || !state))) {
    return PRX_ERR_BAD_ARG;
  }
  const int N = h->ops->num_species;
  if (ix->nvar < ix->ufs + N) {
    return fail(h, PRX_ERR_BAD_ARG, "prx_compute_temp_boxes: state index table does not fit NUM_SPECIES");
  }
  DeviceGuard g(h->device);
  NvtxRange nvtx("prx::computeTemp");
  for (int b = 0; b < nboxes; b++) {
    BoxArgs a;
    std::memset(&a, 0, sizeof(a));
    a.layout = LAYOUT_STATE;
    long long nc = 1;
    for (int d = 0; d < 3; d++) {
      a.len[d] = hi[3 * b + d] - lo[3 * b + d] + 1;
      nc *= a.len[d] > 0 ? a.len[d] : 0;
    }
    if (nc == 0) {
      continue;
    }
    if (nc > 2147483647LL || !view_of(state + b, lo + 3 * b, ix->nvar, &a.snew, hi + 3 * b)) {
      return fail(h, PRX_ERR_BAD_ARG, "prx_compute_temp_boxes: null/short Array4 or box outside the array");
    }
    a.ncells = (int)nc;
    a.ix.urho = ix->urho;
    a.ix.ueint = ix->ueint;
    a.ix.utemp = ix->utemp;
    a.ix.ufs = ix->ufs;
    const cudaError_t e = h->ops->compute_temp(a, (cudaStream_t)cuda_stream);
    if (e != cudaSuccess) {
      return fail(h, PRX_ERR_CUDA, "compute_temp", e);
    }
  }
  return PRX_OK;
}

// ---- pc_ctoprim and SimpleTransport::get_transport_coeffs ------------------------------------------------
static bool prim_box(const int* lo, const int* hi, int b, PrimArgs* a)
{
  long long nc = 1;
  for (int d = 0; d < 3; d++) {
    a->len[d] = hi[3 * b + d] - lo[3 * b + d] + 1;
    nc *= a->len[d] > 0 ? a->len[d] : 0;
  }
  a->ncells = (nc > 2147483647LL) ? -1 : (int)nc;
  return nc > 0;
}

int prx_ctoprim_boxes(
  prx_handle* h, int nboxes, const int* lo, const int* hi, const prx_array4* u, const prx_array4* q, const prx_array4* qaux,
  const prx_state_index* ix, const prx_prim_index* qx, double small_num, void* cuda_stream)
{
  if (h == nullptr || nboxes < 0 || ix == nullptr || qx == nullptr || (nboxes > 0 && (!lo || !hi || !u || !q || !qaux))) {
    return PRX_ERR_BAD_ARG;
  }
  const int N = h->ops->num_species;
  if (ix->nvar < ix->ufs + N || qx->qvar < qx->qfs + N || qx->nqaux < 6) {
    return fail(h, PRX_ERR_BAD_ARG, "prx_ctoprim_boxes: index tables do not fit NUM_SPECIES");
  }
  DeviceGuard g(h->device);
  NvtxRange nvtx("prx::ctoprim");
  for (int b = 0; b < nboxes; b++) {
    PrimArgs a;
    std::memset(&a, 0, sizeof(a));
    if (!prim_box(lo, hi, b, &a)) {
      continue;
    }
    if (a.ncells < 0 || !view_of(u + b, lo + 3 * b, ix->nvar, &a.u, hi + 3 * b) || !view_of(q + b, lo + 3 * b, qx->qvar, &a.q, hi + 3 * b) ||
        !view_of(qaux + b, lo + 3 * b, qx->nqaux, &a.qa, hi + 3 * b)) {
      return fail(h, PRX_ERR_BAD_ARG, "prx_ctoprim_boxes: null/short Array4 or box outside the array");
    }
    a.ix.urho = ix->urho;
    a.ix.umx = ix->umx;
    a.ix.umy = ix->umy;
    a.ix.umz = ix->umz;
    a.ix.ueden = ix->ueden;
    a.ix.ueint = ix->ueint;
    a.ix.utemp = ix->utemp;
    a.ix.ufs = ix->ufs;
    a.qx.qrho = qx->qrho;
    a.qx.qu = qx->qu;
    a.qx.qv = qx->qv;
    a.qx.qw = qx->qw;
    a.qx.qgame = qx->qgame;
    a.qx.qpres = qx->qpres;
    a.qx.qreint = qx->qreint;
    a.qx.qtemp = qx->qtemp;
    a.qx.qfs = qx->qfs;
    a.qx.qgamc = qx->qgamc;
    a.qx.qc = qx->qc;
    a.qx.qcsml = qx->qcsml;
    a.qx.qdpdr = qx->qdpdr;
    a.qx.qdpde = qx->qdpde;
    a.qx.qrspec = qx->qrspec;
    a.small_num = small_num;
    const cudaError_t e = h->ops->ctoprim(a, (cudaStream_t)cuda_stream);
    if (e != cudaSuccess) {
      return fail(h, PRX_ERR_CUDA, "ctoprim", e);
    }
  }
  return PRX_OK;
}

int prx_transport_boxes(
  prx_handle* h, int nboxes, const int* lo, const int* hi, const prx_array4* Y, const prx_array4* T, const prx_array4* rho,
  const prx_array4* D, const prx_array4* mu, const prx_array4* xi, const prx_array4* lam, void* cuda_stream)
{
  if (h == nullptr || nboxes < 0 || (nboxes > 0 && (!lo || !hi || !Y || !T || !rho || !D || !mu || !xi || !lam))) {
    return PRX_ERR_BAD_ARG;
  }
  const int N = h->ops->num_species;
  DeviceGuard g(h->device);
  NvtxRange nvtx("prx::get_transport_coeffs");
  for (int b = 0; b < nboxes; b++) {
    PrimArgs a;
    std::memset(&a, 0, sizeof(a));
    if (!prim_box(lo, hi, b, &a)) {
      continue;
    }
    const int* l = lo + 3 * b;
    const int* hh = hi + 3 * b;
    if (a.ncells < 0 || !view_of(Y + b, l, N, &a.Y, hh) || !view_of(T + b, l, 1, &a.T, hh) || !view_of(rho + b, l, 1, &a.rho, hh) ||
        !view_of(D + b, l, N, &a.D, hh) || !view_of(mu + b, l, 1, &a.mu, hh) || !view_of(xi + b, l, 1, &a.xi, hh) ||
        !view_of(lam + b, l, 1, &a.lam, hh)) {
      return fail(h, PRX_ERR_BAD_ARG, "prx_transport_boxes: null/short Array4 or box outside the array");
    }
    const cudaError_t e = h->ops->transport(a, (cudaStream_t)cuda_stream);
    if (e != cudaSuccess) {
      return fail(h, PRX_ERR_CUDA, e == cudaErrorNotSupported ? "transport: no fit tables for this mechanism" : "transport", e);
    }
  }
  return PRX_OK;
}

// ---- multi-box entry points: the MFIter loop of the caller (PC/Source/React.cpp:133-138) in one call ------------------
// PeleC hands the reactor 16^3 - 64^3 boxes; a single small box cannot fill 148 SMs and its latency is the critical path of
// its slowest cell.  The boxes of one call are independent, so they are issued round-robin on internal streams (as AMReX does
// with amrex.max_gpu_streams) and run concurrently; the caller's stream waits for all of them.
static int box_fanout_begin(prx_handle* h, cudaStream_t user, int nboxes, int* nstreams)
{
  DeviceGuard g(h->device);
  int want = env_int("PRX_BOX_STREAMS", 8);
  want = want < 1 ? 1 : (want > 8 ? 8 : want);
  while (h->n_box_streams < want) {
    const int i = h->n_box_streams;
    if (i == 0) {
      PRX_CUDA(h, cudaEventCreateWithFlags(&h->box_events[8], cudaEventDisableTiming));
    }
    PRX_CUDA(h, cudaStreamCreateWithFlags(&h->box_streams[i], cudaStreamNonBlocking));
    PRX_CUDA(h, cudaEventCreateWithFlags(&h->box_events[i], cudaEventDisableTiming));
    h->n_box_streams++;
  }
  *nstreams = nboxes < want ? nboxes : want;
  PRX_CUDA(h, cudaEventRecord(h->box_events[8], user)); // the boxes start after the work already queued on the caller's stream
  for (int i = 0; i < *nstreams; i++) {
    PRX_CUDA(h, cudaStreamWaitEvent(h->box_streams[i], h->box_events[8], 0));
  }
  return PRX_OK;
}

static int box_fanout_end(prx_handle* h, cudaStream_t user, int nstreams)
{
  DeviceGuard g(h->device);
  for (int i = 0; i < nstreams; i++) {
    PRX_CUDA(h, cudaEventRecord(h->box_events[i], h->box_streams[i]));
    PRX_CUDA(h, cudaStreamWaitEvent(user, h->box_events[i], 0));
  }
  return PRX_OK;
}

int prx_react_boxes(
  prx_handle* h, int nboxes, const int* lo, const int* hi, const prx_array4* rY, const prx_array4* rYsrc, const prx_array4* T,
  const prx_array4* rE, const prx_array4* rEsrc, const prx_array4* FC, const prx_array4i* mask, double dt, double time,
  void* cuda_stream)
{
  if (h == nullptr || nboxes < 0 || (nboxes > 0 && (!lo || !hi || !rY || !rYsrc || !T || !rE || !rEsrc || !FC))) {
    return PRX_ERR_BAD_ARG;
  }
  if (nboxes == 0) {
    return PRX_OK;
  }
  std::lock_guard<std::mutex> lock(h->host_mutex); // the internal streams / events belong to the handle
  if (h->cfg.reactor_kind == PRX_REACTOR_ARKODE) {
    // box-coupled integrations whose step-size controller carries its history from one call to the next (prx_arkode.cuh):
    // the boxes run one after the other, in the caller's order, on the caller's stream
    for (int b = 0; b < nboxes; b++) {
      const int rcb = prx_react_box(h, lo + 3 * b, hi + 3 * b, rY + b, rYsrc + b, T + b, rE + b, rEsrc + b, FC + b,
                                    mask != nullptr ? mask + b : nullptr, dt, time, cuda_stream);
      if (rcb != PRX_OK) {
        return rcb;
      }
    }
    return PRX_OK;
  }
  int ns = 1;
  {
    // the per-call counters are reset once, on the caller's stream, BEFORE the fan-out event: every box accumulates into them
    DeviceGuard g(h->device);
    PRX_CUDA(h, cudaMemsetAsync(&h->d_stats->v[ST_LAST_STEPS], 0, 2 * sizeof(unsigned long long), (cudaStream_t)cuda_stream));
    h->last_ncells = 0;
  }
  int rc = box_fanout_begin(h, (cudaStream_t)cuda_stream, nboxes, &ns);
  bool keep = true;
  for (int b = 0; b < nboxes && rc == PRX_OK; b++) {
    h->keep_last_steps = keep;
    rc = prx_react_box(h, lo + 3 * b, hi + 3 * b, rY + b, rYsrc + b, T + b, rE + b, rEsrc + b, FC + b,
                       mask != nullptr ? mask + b : nullptr, dt, time, h->box_streams[b % ns]);
    keep = true;
  }
  h->keep_last_steps = false;
  const int rc2 = box_fanout_end(h, (cudaStream_t)cuda_stream, ns);
  return rc != PRX_OK ? rc : rc2;
}

int prx_react_state_boxes(
  prx_handle* h, int nboxes, const int* lo, const int* hi, const prx_array4* sold, const prx_array4* snew,
  const prx_array4* nonrs, const prx_array4* I_R, const prx_array4* work, const prx_state_index* ix, double dt, int do_update,
  void* cuda_stream)
{
  if (h == nullptr || nboxes < 0 || ix == nullptr || (nboxes > 0 && (!lo || !hi || !sold || !snew || !nonrs || !I_R))) {
    return PRX_ERR_BAD_ARG;
  }
  if (nboxes == 0) {
    return PRX_OK;
  }
  std::lock_guard<std::mutex> lock(h->host_mutex);
  int ns = 1;
  {
    // the per-call counters are reset once, on the caller's stream, BEFORE the fan-out event: every box accumulates into them
    DeviceGuard g(h->device);
    PRX_CUDA(h, cudaMemsetAsync(&h->d_stats->v[ST_LAST_STEPS], 0, 2 * sizeof(unsigned long long), (cudaStream_t)cuda_stream));
    h->last_ncells = 0;
  }
  int rc = box_fanout_begin(h, (cudaStream_t)cuda_stream, nboxes, &ns);
  bool keep = true;
  for (int b = 0; b < nboxes && rc == PRX_OK; b++) {
    h->keep_last_steps = keep;
    rc = react_state_one(h, lo + 3 * b, hi + 3 * b, sold + b, snew + b, nonrs + b, I_R + b, work != nullptr ? work + b : nullptr,
                         ix, dt, do_update, h->box_streams[b % ns]);
    keep = true;
  }
  h->keep_last_steps = false;
  const int rc2 = box_fanout_end(h, (cudaStream_t)cuda_stream, ns);
  return rc != PRX_OK ? rc : rc2;
}

int prx_react_flat(
  prx_handle* h, double* y, const double* ysrc, double* e, const double* esrc, double* fc, double dt,
  double time, int ncells, void* cuda_stream)
{
  if (h == nullptr || ncells < 0) {
    return PRX_ERR_BAD_ARG;
  }
  if (ncells == 0) {
    return PRX_OK;
  }
  if (y == nullptr || ysrc == nullptr || e == nullptr || esrc == nullptr) {
    return fail(h, PRX_ERR_BAD_ARG, "prx_react_flat: null pointer");
  }
  BoxArgs a;
  std::memset(&a, 0, sizeof(a));
  a.layout = LAYOUT_FLAT;
  a.ncells = ncells;
  a.len[0] = ncells;
  a.len[1] = a.len[2] = 1;
  a.y = y;
  a.ysrc = ysrc;
  a.e = e;
  a.esrc = esrc;
  a.fc = fc;
  a.dt = dt;
  a.time = time;
  if (h->cfg.reactor_kind == PRX_REACTOR_ARKODE) {
    return react_arkode(h, a, (cudaStream_t)cuda_stream); // ReactorArkode.cpp:266-380; prx_last_react_return() = nfe
  }
  return react_args(h, a, (cudaStream_t)cuda_stream);
}

int prx_react_box_host(
  prx_handle* h, int nx, int ny, int nz, double* rY, const double* rYsrc, double* T, double* rE,
  const double* rEsrc, double* FC, const int* mask, double dt, double time)
{
  if (h == nullptr || nx < 0 || ny < 0 || nz < 0) {
    return PRX_ERR_BAD_ARG;
  }
  const long long nc = (long long)nx * ny * nz;
  if (nc == 0) {
    return PRX_OK;
  }
  if (rY == nullptr || rYsrc == nullptr || T == nullptr || rE == nullptr || rEsrc == nullptr || FC == nullptr) {
    return fail(h, PRX_ERR_BAD_ARG, "prx_react_box_host: null pointer");
  }
  std::lock_guard<std::mutex> host_lock(h->host_mutex);
  DeviceGuard g(h->device);
  const int N = h->ops->num_species;
  // device staging: [rY N][rYsrc N][T][rE][rEsrc][FC] doubles + mask ints
  const size_t nd = (size_t)(2 * N + 4) * nc;
  const size_t bytes = nd * sizeof(double) + (size_t)nc * sizeof(int);
  const int rc = ensure_stage(h, bytes);
  if (rc != PRX_OK) {
    return rc;
  }
  cudaStream_t st = h->stage_stream;
  double* d_rY = h->d_stage;
  double* d_rYsrc = d_rY + (size_t)N * nc;
  double* d_T = d_rYsrc + (size_t)N * nc;
  double* d_rE = d_T + nc;
  double* d_rEsrc = d_rE + nc;
  double* d_FC = d_rEsrc + nc;
  int* d_mask = reinterpret_cast<int*>(d_FC + nc);
  // Slab pipeline: the box is cut along k into slabs (PRX_HOST_SLABS, default 8); slab s+1 is copied in while slab s reacts
  // and slab s-1 is copied out (copy-in, copy-out and PRX_HOST_REACT_STREAMS react streams, full-duplex PCIe).  Measured
  // (r02ae, 128^3 LiDryer, device-timed 67.1 ms): one react stream 2 slabs 75.6 ms / 4 slabs 82.3 ms (every slab's persistent
  // kernel drains alone); two react streams 2 / 4 / 8 / 12 slabs 73.2 / 71.8 / 70.9 / 70.5 ms.  Every array is SoA with component stride nc, so one slab
  // of one array is a 2-D copy: rows = components (pitch nc), row length = the slab's cells.
  const long long plane = (long long)nx * ny;
  int nslab = (nz >= 8 && nc >= (1LL << 18)) ? env_int("PRX_HOST_SLABS", 8) : 1;
  nslab = nslab < 1 ? 1 : (nslab > 14 ? 14 : nslab);
  nslab = nslab > nz ? nz : nslab;
  // Slab boundaries along k.  What the pipeline cannot hide is the H2D of the FIRST slab and the D2H of the LAST one, so
  // those two are cut thin (PRX_HOST_EDGE_PLANES k-planes each; default 0 = equal slabs: measured r02s on 128^3 LiDryer, 4 planes 74.8 ms vs 76.7 ms, 8 / 16 planes no gain - the thin slabs run as small, less efficient launches) and the nslab slabs in
  // between share the rest equally.
  int kcut[18];
  int ncut = 0;
  {
    int edge = (nslab >= 2) ? env_int("PRX_HOST_EDGE_PLANES", 0) : 0;
    if (edge < 0 || 2 * edge + nslab > nz) {
      edge = 0;
    }
    kcut[ncut++] = 0;
    if (edge > 0) {
      kcut[ncut++] = edge;
    }
    const int mid0 = edge, mid1 = nz - edge;
    for (int sidx = 1; sidx <= nslab; sidx++) {
      kcut[ncut++] = mid0 + (int)((long long)(mid1 - mid0) * sidx / nslab);
    }
    if (edge > 0) {
      kcut[ncut++] = nz;
    }
    nslab = ncut - 1;
  }
  auto mk = [&](double* p, int ncomp) {
    prx_array4 a;
    a.p = p;
    a.jstride = nx;
    a.kstride = plane;
    a.nstride = nc;
    a.begin[0] = a.begin[1] = a.begin[2] = 0;
    a.end[0] = nx;
    a.end[1] = ny;
    a.end[2] = nz;
    a.ncomp = ncomp;
    return a;
  };
  prx_array4 a_rY = mk(d_rY, N), a_rYsrc = mk(d_rYsrc, N), a_T = mk(d_T, 1), a_rE = mk(d_rE, 1),
             a_rEsrc = mk(d_rEsrc, 1), a_FC = mk(d_FC, 1);
  prx_array4i a_mask;
  a_mask.p = (mask != nullptr) ? d_mask : nullptr;
  a_mask.jstride = nx;
  a_mask.kstride = plane;
  a_mask.nstride = nc;
  a_mask.begin[0] = a_mask.begin[1] = a_mask.begin[2] = 0;
  a_mask.end[0] = nx;
  a_mask.end[1] = ny;
  a_mask.end[2] = nz;
  a_mask.ncomp = 1;
  cudaEvent_t ev_in[16], ev_done[16];
  for (int sidx = 0; sidx < nslab; sidx++) {
    PRX_CUDA(h, cudaEventCreateWithFlags(&ev_in[sidx], cudaEventDisableTiming));
    PRX_CUDA(h, cudaEventCreateWithFlags(&ev_done[sidx], cudaEventDisableTiming));
  }
  auto copy2d = [&](void* dst, const void* src, size_t elem, long long off, long long cells, int ncomp, cudaMemcpyKind kind,
                    cudaStream_t strm) {
    return cudaMemcpy2DAsync(
      (char*)dst + off * elem, (size_t)nc * elem, (const char*)src + off * elem, (size_t)nc * elem, (size_t)cells * elem,
      (size_t)ncomp, kind, strm);
  };
  int rc_react = PRX_OK;
  // an error inside the pipeline must not return before the three streams are drained and the events destroyed
  // (the async copies still touch the caller's host buffers): record it and fall through to the common exit
#define SLAB_CUDA(call)                                        \
  do {                                                         \
    const cudaError_t e_ = (call);                             \
    if (e_ != cudaSuccess && rc_react == PRX_OK) {             \
      rc_react = fail(h, PRX_ERR_CUDA, #call, e_);             \
    }                                                          \
  } while (0)
  // React streams: slab s reacts on stream s mod nrs.  The integrator kernels are persistent (blocks leave only when the
  // work queue is empty and their last cells have finished), so on ONE stream every slab pays its own drain; on alternating
  // streams the next slab's blocks move in as the previous slab's leave, and thinner slabs (less exposed first-slab H2D /
  // last-slab D2H) stop costing kernel efficiency.  PRX_HOST_REACT_STREAMS (1-4, default 2).
  int nrs = env_int("PRX_HOST_REACT_STREAMS", 2);
  nrs = nrs < 1 ? 1 : (nrs > 4 ? 4 : nrs);
  nrs = nrs > nslab ? nslab : nrs;
  cudaStream_t rs[4] = {st, nullptr, nullptr, nullptr};
  for (int i = 1; i < nrs; i++) {
    if (h->host_react_streams[i - 1] == nullptr) {
      SLAB_CUDA(cudaStreamCreateWithFlags(&h->host_react_streams[i - 1], cudaStreamNonBlocking));
    }
    rs[i] = h->host_react_streams[i - 1];
  }
  // the per-call step sum is reset once, on the first react stream, before any slab's kernel may add to it
  h->keep_last_steps = true;
  h->last_ncells = 0;
  cudaEvent_t ev_reset = nullptr;
  SLAB_CUDA(cudaMemsetAsync(&h->d_stats->v[ST_LAST_STEPS], 0, 2 * sizeof(unsigned long long), st));
  if (nrs > 1) {
    SLAB_CUDA(cudaEventCreateWithFlags(&ev_reset, cudaEventDisableTiming));
    SLAB_CUDA(cudaEventRecord(ev_reset, st));
    for (int i = 1; i < nrs && rc_react == PRX_OK; i++) {
      SLAB_CUDA(cudaStreamWaitEvent(rs[i], ev_reset, 0));
    }
  }
  for (int sidx = 0; sidx < nslab && rc_react == PRX_OK; sidx++) {
    cudaStream_t st = rs[sidx % nrs];
    const int k0 = kcut[sidx], k1 = kcut[sidx + 1];
    const long long off = plane * k0, cells = plane * (k1 - k0);
    SLAB_CUDA(copy2d(d_rY, rY, sizeof(double), off, cells, N, cudaMemcpyHostToDevice, h->in_stream));
    SLAB_CUDA(copy2d(d_rYsrc, rYsrc, sizeof(double), off, cells, N, cudaMemcpyHostToDevice, h->in_stream));
    SLAB_CUDA(copy2d(d_T, T, sizeof(double), off, cells, 1, cudaMemcpyHostToDevice, h->in_stream));
    SLAB_CUDA(copy2d(d_rE, rE, sizeof(double), off, cells, 1, cudaMemcpyHostToDevice, h->in_stream));
    SLAB_CUDA(copy2d(d_rEsrc, rEsrc, sizeof(double), off, cells, 1, cudaMemcpyHostToDevice, h->in_stream));
    if (mask != nullptr) {
      SLAB_CUDA(copy2d(d_mask, mask, sizeof(int), off, cells, 1, cudaMemcpyHostToDevice, h->in_stream));
    }
    SLAB_CUDA(cudaEventRecord(ev_in[sidx], h->in_stream));
    SLAB_CUDA(cudaStreamWaitEvent(st, ev_in[sidx], 0));
    const int lo[3] = {0, 0, k0};
    const int hi[3] = {nx - 1, ny - 1, k1 - 1};
    rc_react = prx_react_box(h, lo, hi, &a_rY, &a_rYsrc, &a_T, &a_rE, &a_rEsrc, &a_FC, &a_mask, dt, time, st);
    if (rc_react != PRX_OK) {
      break;
    }
    SLAB_CUDA(cudaEventRecord(ev_done[sidx], st));
    SLAB_CUDA(cudaStreamWaitEvent(h->out_stream, ev_done[sidx], 0));
    SLAB_CUDA(copy2d(rY, d_rY, sizeof(double), off, cells, N, cudaMemcpyDeviceToHost, h->out_stream));
    SLAB_CUDA(copy2d(T, d_T, sizeof(double), off, cells, 1, cudaMemcpyDeviceToHost, h->out_stream));
    SLAB_CUDA(copy2d(rE, d_rE, sizeof(double), off, cells, 1, cudaMemcpyDeviceToHost, h->out_stream));
    SLAB_CUDA(copy2d(FC, d_FC, sizeof(double), off, cells, 1, cudaMemcpyDeviceToHost, h->out_stream));
  }
#undef SLAB_CUDA
  h->keep_last_steps = false;
  cudaError_t es = cudaStreamSynchronize(h->out_stream);
  cudaError_t es2 = cudaStreamSynchronize(st);
  for (int i = 1; i < nrs; i++) {
    if (rs[i] != nullptr) {
      const cudaError_t er = cudaStreamSynchronize(rs[i]);
      es2 = (es2 == cudaSuccess) ? er : es2;
    }
  }
  cudaError_t es3 = cudaStreamSynchronize(h->in_stream);
  if (ev_reset != nullptr) {
    cudaEventDestroy(ev_reset);
  }
  for (int sidx = 0; sidx < nslab; sidx++) {
    cudaEventDestroy(ev_in[sidx]);
    cudaEventDestroy(ev_done[sidx]);
  }
  if (rc_react != PRX_OK) {
    return rc_react;
  }
  PRX_CUDA(h, es);
  PRX_CUDA(h, es2);
  PRX_CUDA(h, es3);
  return PRX_OK;
}

static int box_args(
  prx_handle* h, const char* who, const int lo[3], const int hi[3], const prx_array4* rY, const prx_array4* rYsrc,
  const prx_array4* T, const prx_array4* rE, const prx_array4* rEsrc, const prx_array4* FC, BoxArgs* out, bool* empty)
{
  const int N = h->ops->num_species;
  BoxArgs& a = *out;
  std::memset(&a, 0, sizeof(a));
  a.layout = LAYOUT_ARRAY4;
  long long nc = 1;
  for (int d = 0; d < 3; d++) {
    a.len[d] = hi[d] - lo[d] + 1;
    if (a.len[d] <= 0) {
      a.len[d] = 0;
    }
    nc *= a.len[d];
  }
  *empty = (nc == 0);
  if (nc == 0) {
    return PRX_OK;
  }
  if (nc > 2147483647LL) {
    return fail(h, PRX_ERR_BAD_ARG, "box too large (int32 cell index)");
  }
  a.ncells = (int)nc;
  if ((rY && !view_of(rY, lo, N, &a.rY)) || (rYsrc && !view_of(rYsrc, lo, N, &a.rYsrc)) || (T && !view_of(T, lo, 1, &a.T)) ||
      (rE && !view_of(rE, lo, 1, &a.rE)) || (rEsrc && !view_of(rEsrc, lo, 1, &a.rEsrc)) || (FC && !view_of(FC, lo, 1, &a.FC))) {
    return fail(h, PRX_ERR_BAD_ARG, who);
  }
  return PRX_OK;
}

int prx_flatten(
  prx_handle* h, const int lo[3], const int hi[3], int ncells, int ordering, const prx_array4* rY,
  const prx_array4* rYsrc, const prx_array4* T, const prx_array4* rE, const prx_array4* rEsrc, double* y,
  double* ysrc, double* e, double* esrc, void* cuda_stream)
{
  if (h == nullptr || lo == nullptr || hi == nullptr) {
    return PRX_ERR_BAD_ARG;
  }
  if (!rY || !rYsrc || !T || !rE || !rEsrc || !y || !ysrc || !e || !esrc || (ordering != 0 && ordering != 1)) {
    return fail(h, PRX_ERR_BAD_ARG, "prx_flatten: null pointer or bad ordering");
  }
  BoxArgs a;
  bool empty = false;
  const int rc = box_args(h, "prx_flatten: short Array4 or box outside the array", lo, hi, rY, rYsrc, T, rE, rEsrc, nullptr, &a, &empty);
  if (rc != PRX_OK || empty) {
    return rc;
  }
  if (ncells != a.ncells) {
    return fail(h, PRX_ERR_BAD_ARG, "prx_flatten: ncells != box.numPts()");
  }
  DeviceGuard g(h->device);
  PRX_CUDA(h, h->ops->flatten(a, h->cfg, ordering, y, ysrc, e, esrc, (cudaStream_t)cuda_stream));
  return PRX_OK;
}

int prx_unflatten(
  prx_handle* h, const int lo[3], const int hi[3], int ncells, int ordering, const prx_array4* rY,
  const prx_array4* T, const prx_array4* rE, const prx_array4* rEsrc, const prx_array4* FC, const double* y,
  const double* e, const long long* fcunt, double dt, void* cuda_stream)
{
  if (h == nullptr || lo == nullptr || hi == nullptr) {
    return PRX_ERR_BAD_ARG;
  }
  if (!rY || !T || !rE || !rEsrc || !FC || !y || !e || !fcunt || (ordering != 0 && ordering != 1)) {
    return fail(h, PRX_ERR_BAD_ARG, "prx_unflatten: null pointer or bad ordering");
  }
  BoxArgs a;
  bool empty = false;
  const int rc = box_args(h, "prx_unflatten: short Array4 or box outside the array", lo, hi, rY, nullptr, T, rE, rEsrc, FC, &a, &empty);
  if (rc != PRX_OK || empty) {
    return rc;
  }
  if (ncells != a.ncells) {
    return fail(h, PRX_ERR_BAD_ARG, "prx_unflatten: ncells != box.numPts()");
  }
  a.dt = dt;
  DeviceGuard g(h->device);
  PRX_CUDA(h, h->ops->unflatten(a, h->cfg, ordering, y, e, fcunt, (cudaStream_t)cuda_stream));
  return PRX_OK;
}

int prx_react_flat_host(
  prx_handle* h, double* y, const double* ysrc, double* e, const double* esrc, double* fc, double dt, double time,
  int ncells)
{
  if (h == nullptr || ncells < 0) {
    return PRX_ERR_BAD_ARG;
  }
  if (ncells == 0) {
    return PRX_OK;
  }
  std::lock_guard<std::mutex> host_lock(h->host_mutex);
  if (y == nullptr || ysrc == nullptr || e == nullptr || esrc == nullptr) {
    return fail(h, PRX_ERR_BAD_ARG, "prx_react_flat_host: null pointer");
  }
  DeviceGuard g(h->device);
  const int N = h->ops->num_species;
  const size_t nc = (size_t)ncells;
  const size_t nd = nc * (size_t)(2 * N + 4);
  const int rc = ensure_stage(h, nd * sizeof(double));
  if (rc != PRX_OK) {
    return rc;
  }
  cudaStream_t st = h->stage_stream;
  double* d_y = h->d_stage;
  double* d_ysrc = d_y + nc * (N + 1);
  double* d_e = d_ysrc + nc * N;
  double* d_esrc = d_e + nc;
  double* d_fc = d_esrc + nc;
  PRX_CUDA(h, cudaMemcpyAsync(d_y, y, sizeof(double) * nc * (N + 1), cudaMemcpyHostToDevice, st));
  PRX_CUDA(h, cudaMemcpyAsync(d_ysrc, ysrc, sizeof(double) * nc * N, cudaMemcpyHostToDevice, st));
  PRX_CUDA(h, cudaMemcpyAsync(d_e, e, sizeof(double) * nc, cudaMemcpyHostToDevice, st));
  PRX_CUDA(h, cudaMemcpyAsync(d_esrc, esrc, sizeof(double) * nc, cudaMemcpyHostToDevice, st));
  const int r = prx_react_flat(h, d_y, d_ysrc, d_e, d_esrc, d_fc, dt, time, ncells, st);
  if (r != PRX_OK) {
    return r;
  }
  PRX_CUDA(h, cudaMemcpyAsync(y, d_y, sizeof(double) * nc * (N + 1), cudaMemcpyDeviceToHost, st));
  PRX_CUDA(h, cudaMemcpyAsync(e, d_e, sizeof(double) * nc, cudaMemcpyDeviceToHost, st));
  if (fc != nullptr) {
    PRX_CUDA(h, cudaMemcpyAsync(fc, d_fc, sizeof(double) * nc, cudaMemcpyDeviceToHost, st));
  }
  PRX_CUDA(h, cudaStreamSynchronize(st));
  return PRX_OK;
}

int prx_find_mechanism(int num_species, int num_reactions, const char* first_species, const char* last_species)
{
  for (int m = 0; m < g_num_mechs; m++) {
    const MechOps* o = ops_for(m);
    if (o == nullptr || o->num_species != num_species || o->num_reactions != num_reactions) {
      continue;
    }
    if (first_species != nullptr && std::strcmp(first_species, o->species_names[0]) != 0) {
      continue;
    }
    if (last_species != nullptr && std::strcmp(last_species, o->species_names[num_species - 1]) != 0) {
      continue;
    }
    return m;
  }
  return PRX_ERR_UNSUPPORTED;
}

int prx_num_mechanisms() { return g_num_mechs; }

int prx_mechanism_id(const char* name)
{
  for (int m = 0; name != nullptr && m < g_num_mechs; m++) {
    if (std::strcmp(name, ops_for(m)->name) == 0) {
      return m;
    }
  }
  return PRX_ERR_UNSUPPORTED;
}

const char* prx_species_name(int mechanism, int k)
{
  const MechOps* o = ops_for(mechanism);
  return (o != nullptr && k >= 0 && k < o->num_species) ? o->species_names[k] : nullptr;
}

int prx_reset_stats(prx_handle* h)
{
  if (h == nullptr) {
    return PRX_ERR_BAD_ARG;
  }
  DeviceGuard g(h->device);
  PRX_CUDA(h, cudaDeviceSynchronize());
  PRX_CUDA(h, cudaMemset(h->d_stats, 0, sizeof(DevStats)));
  h->ark_nfe = h->ark_nst = h->ark_netf = h->ark_too_much_work = h->ark_cells = 0;
  return PRX_OK;
}

int prx_get_stats(prx_handle* h, prx_stats* out)
{
  if (h == nullptr || out == nullptr) {
    return PRX_ERR_BAD_ARG;
  }
  DeviceGuard g(h->device);
  PRX_CUDA(h, cudaDeviceSynchronize());
  DevStats s;
  PRX_CUDA(h, cudaMemcpy(&s, h->d_stats, sizeof(s), cudaMemcpyDeviceToHost));
  out->cells = (long long)s.v[ST_CELLS];
  out->cells_masked = (long long)s.v[ST_MASKED];
  out->n_steps = (long long)s.v[ST_STEPS];
  out->n_rhs = (long long)s.v[ST_RHS];
  out->n_jac = (long long)s.v[ST_JAC];
  out->n_setups = (long long)s.v[ST_SETUPS];
  out->n_newton = (long long)s.v[ST_NEWTON];
  out->n_errfail = (long long)s.v[ST_ERRFAIL];
  out->n_convfail = (long long)s.v[ST_CONVFAIL];
  out->n_failed = (long long)s.v[ST_FAILED];
  out->max_steps = (long long)s.v[ST_MAXSTEPS];
  out->reserved = 0;
  if (h->cfg.reactor_kind == PRX_REACTOR_ARKODE) { // box-coupled: steps / RHS sweeps / error-test failures of the boxes, summed
    out->cells = h->ark_cells;
    out->n_steps = h->ark_nst;
    out->n_rhs = h->ark_nfe;
    out->n_errfail = h->ark_netf;
    out->n_failed = h->ark_too_much_work;
  }
  return PRX_OK;
}

const char* prx_kernel_name(const prx_handle* h)
{
  return h ? h->ops->kernel_name(h->cfg.reactor_kind, h->sched) : "";
}

int prx_last_react_return(prx_handle* h)
{
  if (h == nullptr) {
    return PRX_ERR_BAD_ARG;
  }
  if (h->cfg.reactor_kind == PRX_REACTOR_CVODE && h->last_layout == LAYOUT_ARRAY4) {
    return 20; // ReactorCvode.cpp:1463-1464: "Dummy, the return value is no longer used"
  }
  if (h->cfg.reactor_kind == PRX_REACTOR_ARKODE) {
    return (int)h->ark_last_nfe; // ReactorArkode.cpp:259: return static_cast<int>(nfe)
  }
  DeviceGuard g(h->device);
  if (cudaDeviceSynchronize() != cudaSuccess || h->last_ncells <= 0) {
    return 0;
  }
  unsigned long long v = 0;
  if (h->cfg.reactor_kind == PRX_REACTOR_CVODE) {
    // 1-D react (ReactorCvode.cpp:1673-1687): nfe + nfeLS of the call.  The reference integrates the ncells cells of one
    // call as ONE coupled system; here every cell is its own system (see include/pelerx_b200.h), so the work count
    // reported is the largest per-cell nfe + nfeDQ - identical to the reference's for ncells = 1.
    if (cudaMemcpy(&v, &h->d_stats->v[ST_LAST_MAXRHS], sizeof(v), cudaMemcpyDeviceToHost) != cudaSuccess) {
      return 0;
    }
    return (int)v;
  }
  if (cudaMemcpy(&v, &h->d_stats->v[ST_LAST_STEPS], sizeof(v), cudaMemcpyDeviceToHost) != cudaSuccess) {
    return 0;
  }
  return (int)((double)v / (double)h->last_ncells); // ReactorRK64.cpp:316-319
}

int prx_probe_rhs_host(
  prx_handle* h, int n, double t, const double* y, const double* rhoe, const double* rhoesrc,
  const double* rYsrc, double* ydot)
{
  if (h == nullptr || n <= 0 || !y || !rhoe || !rhoesrc || !rYsrc || !ydot) {
    return PRX_ERR_BAD_ARG;
  }
  std::lock_guard<std::mutex> host_lock(h->host_mutex);
  DeviceGuard g(h->device);
  const int N = h->ops->num_species, NEQ = N + 1;
  const size_t nd = (size_t)n * (2 * NEQ + N + 2);
  const int rc = ensure_stage(h, nd * sizeof(double));
  if (rc != PRX_OK) {
    return rc;
  }
  double* d_y = h->d_stage;
  double* d_yd = d_y + (size_t)n * NEQ;
  double* d_src = d_yd + (size_t)n * NEQ;
  double* d_e = d_src + (size_t)n * N;
  double* d_es = d_e + n;
  cudaStream_t st = h->stage_stream;
  PRX_CUDA(h, cudaMemcpyAsync(d_y, y, sizeof(double) * n * NEQ, cudaMemcpyHostToDevice, st));
  PRX_CUDA(h, cudaMemcpyAsync(d_src, rYsrc, sizeof(double) * n * N, cudaMemcpyHostToDevice, st));
  PRX_CUDA(h, cudaMemcpyAsync(d_e, rhoe, sizeof(double) * n, cudaMemcpyHostToDevice, st));
  PRX_CUDA(h, cudaMemcpyAsync(d_es, rhoesrc, sizeof(double) * n, cudaMemcpyHostToDevice, st));
  PRX_CUDA(h, h->ops->probe_rhs(n, t, h->cfg.reactor_type, d_y, d_e, d_es, d_src, d_yd, st));
  PRX_CUDA(h, cudaMemcpyAsync(ydot, d_yd, sizeof(double) * n * NEQ, cudaMemcpyDeviceToHost, st));
  PRX_CUDA(h, cudaStreamSynchronize(st));
  return PRX_OK;
}

int prx_probe_jac_host(prx_handle* h, int n, const double* y, double* J)
{
  if (h == nullptr || n <= 0 || !y || !J) {
    return PRX_ERR_BAD_ARG;
  }
  std::lock_guard<std::mutex> host_lock(h->host_mutex);
  DeviceGuard g(h->device);
  const int N = h->ops->num_species, NEQ = N + 1;
  const size_t nd = (size_t)n * (NEQ + (size_t)NEQ * NEQ);
  const int rc = ensure_stage(h, nd * sizeof(double));
  if (rc != PRX_OK) {
    return rc;
  }
  double* d_y = h->d_stage;
  double* d_J = d_y + (size_t)n * NEQ;
  cudaStream_t st = h->stage_stream;
  PRX_CUDA(h, cudaMemcpyAsync(d_y, y, sizeof(double) * n * NEQ, cudaMemcpyHostToDevice, st));
  PRX_CUDA(h, h->ops->probe_jac(n, h->cfg.reactor_type, d_y, d_J, st));
  PRX_CUDA(h, cudaMemcpyAsync(J, d_J, sizeof(double) * n * NEQ * NEQ, cudaMemcpyDeviceToHost, st));
  PRX_CUDA(h, cudaStreamSynchronize(st));
  return PRX_OK;
}

} // extern "C"
