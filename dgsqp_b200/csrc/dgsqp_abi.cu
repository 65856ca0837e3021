// C-ABI entry points (include/dgsqp_b200.h) and the persistent solve kernel.
//
// One CTA solves one game instance at a time and pulls the next instance index from a global
// atomic counter (iteration counts range 3..50+ so static assignment would idle most of the grid).
// Instances are independent: no inter-CTA communication on the solve path.
#include <cuda_runtime.h>
#include <atomic>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <string>
#include <new>

#include "sqp_v2.cuh"
#include "host_setup.h"

#define DGSQP_VERSION_STR "dgsqp_b200 0.1.0 (sm_100a)"

static thread_local std::string g_last_error;
static std::atomic<long long> g_launches{0};

static int set_err(int code, const std::string& msg) { g_last_error = msg; return code; }
#define CUDA_TRY(expr)                                                                           \
  do {                                                                                           \
    cudaError_t e__ = (expr);                                                                    \
    if (e__ != cudaSuccess)                                                                      \
      return set_err(DGSQP_ECUDA, std::string(#expr) + ": " + cudaGetErrorString(e__));          \
  } while (0)

struct KernelArgs {
  int B;
  const double* x0; const double* u_ws; const double* l_ws;
  double* u_out; double* l_out; double* x_out; double* cost_out; double* cond_out;
  int* num_iters; int* status; int* qp_solves; int* diag; long long* phase;
  double* ws; size_t ws_stride; size_t smem_doubles; size_t smem_used;
  int* counter;
  int poison;     // debug (DGSQP_POISON=1): NaN-fill the CTA's whole workspace before every instance
};

// SM = true: every hot buffer of the memory plan is shared-memory resident (plan.hot_in_smem)
template <bool SM>
__global__ void __launch_bounds__(DG_MAX_THREADS, 1) dgsqp_solve_kernel(const GameDesc* __restrict__ Gp, const SolverParams* __restrict__ Pp, KernelArgs A) {
  __shared__ GameDesc sG;
  __shared__ SolverParams sP;
  extern __shared__ double s_dyn[];
  __shared__ int s_inst;
  {
    const int nw = (int)(sizeof(GameDesc) / sizeof(int));
    for (int i = threadIdx.x; i < nw; i += blockDim.x) ((int*)&sG)[i] = ((const int*)Gp)[i];
    const int np = (int)(sizeof(SolverParams) / sizeof(int));
    for (int i = threadIdx.x; i < np; i += blockDim.x) ((int*)&sP)[i] = ((const int*)Pp)[i];
  }
  __syncthreads();
  Cta c;
  c.flip = 0;
  // the solve context (dimensions + the table of buffer pointers) lives in shared memory: in a per-thread
  // local-memory copy every buffer access of the solver would start with a local load
  __shared__ SolveCtx sX;
  if (threadIdx.x == 0) {
    sX.G = &sG; sX.P = &sP; sX.D = make_dims(sG.M, sG.N);
    plan_memory(sX.D, A.ws + (size_t)blockIdx.x * A.ws_stride, s_dyn, A.smem_doubles, sX.W);
  }
  __syncthreads();
  SolveCtx& X = sX;
  const Dims& D = X.D;
  while (true) {
    if (threadIdx.x == 0) s_inst = atomicAdd(A.counter, 1);
    __syncthreads();
    const int inst = s_inst;
    __syncthreads();
    if (inst >= A.B) break;
    if (A.poison) {
      const double qnan = __longlong_as_double(0x7ff8dead0000beefLL);
      for (size_t i = threadIdx.x; i < A.smem_used; i += blockDim.x) s_dyn[i] = qnan;
      double* wsl = A.ws + (size_t)blockIdx.x * A.ws_stride;
      for (size_t i = threadIdx.x; i < A.ws_stride; i += blockDim.x) wsl[i] = qnan;
      __syncthreads();
    }
    if (threadIdx.x == 0) X.x0 = A.x0 + (size_t)inst * D.nq;
    SolveOut O;
    O.u = A.u_out + (size_t)inst * D.n;
    O.l = A.l_out + (size_t)inst * D.m;
    O.x = A.x_out + (size_t)inst * (D.N + 1) * D.nq;
    O.cost = A.cost_out + (size_t)inst * D.M;
    O.cond = A.cond_out + (size_t)inst * 3;
    O.num_iters = A.num_iters + inst; O.status = A.status + inst; O.qp_solves = A.qp_solves + inst;
    O.diag = A.diag ? A.diag + (size_t)inst * DG_NDIAG : nullptr;
    O.l_init = nullptr;
    if (threadIdx.x == 0) { for (int i = 0; i < DG_NPHASE; ++i) dg_s_ph[i] = 0; dg_s_ph[DG_NPHASE] = clock64(); }
    if (sP.policy == 2) sqp_solve_v2<SM>(c, X, A.u_ws + (size_t)inst * D.n, A.l_ws ? A.l_ws + (size_t)inst * D.m : nullptr, O);
    else sqp_solve_v1<SM>(c, X, A.u_ws + (size_t)inst * D.n, A.l_ws ? A.l_ws + (size_t)inst * D.m : nullptr, O);
    c.lap(PH_OTHER);
    if (threadIdx.x == 0 && A.phase) for (int i = 0; i < DG_NPHASE; ++i) A.phase[(size_t)inst * DG_NPHASE + i] = dg_s_ph[i];
  }
}

// FP64 FMA throughput probe (roofline denominator: MEASURED_PEAKS.json carries no FP64 figure).
__global__ void dgsqp_fp64_probe_kernel(double* out, int iters, double a, double b) {
  double x0 = threadIdx.x * 1e-3, x1 = x0 + 1.0, x2 = x0 + 2.0, x3 = x0 + 3.0, x4 = x0 + 4.0, x5 = x0 + 5.0, x6 = x0 + 6.0, x7 = x0 + 7.0;
  for (int i = 0; i < iters; ++i) {
    x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
    x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
  }
  double s = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
  if (s == 123.456) out[0] = s;      // never true; keeps the chain alive
}

struct dgsqp_handle {
  GameDesc G; SolverParams P; Dims D;
  int device = 0, sm_count = 0, ctas_per_sm = 0, threads = 256, grid_cap = 0;
  size_t ws_doubles = 0, smem_bytes = 0, smem_budget = 0, smem_limit = 0;   // budget/limit in doubles (0 limit = device maximum)
  MemPlan plan; size_t ws_alloc = 0;
  double* d_ws = nullptr; int* d_counter = nullptr; int* d_diag = nullptr; long long* d_phase = nullptr; size_t diag_cap = 0;
  GameDesc* d_G = nullptr; SolverParams* d_P = nullptr;
  // staging for host-pointer calls
  size_t stage_cap = 0;
  double *s_lws = nullptr;
  double *s_x0 = nullptr, *s_uws = nullptr, *s_u = nullptr, *s_l = nullptr, *s_x = nullptr, *s_cost = nullptr, *s_cond = nullptr;
  int *s_it = nullptr, *s_st = nullptr, *s_qp = nullptr;
};

static void free_stage(dgsqp_handle* h) {
  cudaFree(h->s_lws); h->s_lws = nullptr;
  cudaFree(h->s_x0); cudaFree(h->s_uws); cudaFree(h->s_u); cudaFree(h->s_l); cudaFree(h->s_x); cudaFree(h->s_cost);
  cudaFree(h->s_cond); cudaFree(h->s_it); cudaFree(h->s_st); cudaFree(h->s_qp);
  h->s_x0 = h->s_uws = h->s_u = h->s_l = h->s_x = h->s_cost = h->s_cond = nullptr; h->s_it = h->s_st = h->s_qp = nullptr;
  h->stage_cap = 0;
}

static int ensure_grid(dgsqp_handle* h) {
  int occ = 0;
  {
    // shared-memory budget of one CTA: the opt-in maximum minus the kernel's static shared memory
    int optin = 0;
    CUDA_TRY(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, h->device));
    cudaFuncAttributes fa;
    CUDA_TRY(cudaFuncGetAttributes(&fa, dgsqp_solve_kernel<true>));
    cudaFuncAttributes fb;
    CUDA_TRY(cudaFuncGetAttributes(&fb, dgsqp_solve_kernel<false>));
    if (fb.sharedSizeBytes > fa.sharedSizeBytes) fa.sharedSizeBytes = fb.sharedSizeBytes;
    size_t avail = (size_t)optin > fa.sharedSizeBytes + 64 ? ((size_t)optin - fa.sharedSizeBytes - 64) / sizeof(double) : 0;
    if (h->smem_limit && h->smem_limit < avail) avail = h->smem_limit;
    h->smem_budget = avail;
    Workspace tmp;
    h->plan = plan_memory(h->D, nullptr, nullptr, h->smem_budget, tmp);
    h->ws_doubles = h->plan.gmem;
    h->smem_bytes = sizeof(double) * h->plan.smem;
  }
  if (h->plan.hot_in_smem) {
    CUDA_TRY(cudaFuncSetAttribute(dgsqp_solve_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_bytes));
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, dgsqp_solve_kernel<true>, h->threads, h->smem_bytes));
  } else {
    CUDA_TRY(cudaFuncSetAttribute(dgsqp_solve_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_bytes));
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, dgsqp_solve_kernel<false>, h->threads, h->smem_bytes));
  }
  if (occ < 1) return set_err(DGSQP_ECUDA, "kernel does not fit on an SM");
  int per_sm = h->ctas_per_sm > 0 ? (h->ctas_per_sm < occ ? h->ctas_per_sm : occ) : occ;
  int cap = per_sm * h->sm_count;
  if (cap != h->grid_cap || h->ws_doubles != h->ws_alloc) {
    if (h->d_ws) { cudaFree(h->d_ws); h->d_ws = nullptr; }
    CUDA_TRY(cudaMalloc(&h->d_ws, sizeof(double) * h->ws_doubles * (size_t)cap));
    CUDA_TRY(cudaMemset(h->d_ws, 0, sizeof(double) * h->ws_doubles * (size_t)cap));
    h->grid_cap = cap; h->ws_alloc = h->ws_doubles;
  }
  return 0;
}

extern "C" {

const char* dgsqp_last_error(void) { return g_last_error.c_str(); }
const char* dgsqp_version(void) { return DGSQP_VERSION_STR; }
int64_t dgsqp_kernel_launches(void) { return (int64_t)g_launches.load(); }

static int create_common(const dgsqp_racing_game* game, const dgsqp_params* params, const dgsqp_v2_params* params2, int device, dgsqp_handle** out) {
  if (!out) return set_err(DGSQP_EINVAL, "out is NULL");
  *out = nullptr;
  dgsqp_handle* h = new (std::nothrow) dgsqp_handle();
  if (!h) return set_err(DGSQP_ENOMEM, "host allocation failed");
  if (dg_fill_game(game, &h->G) != 0) { delete h; return set_err(DGSQP_EINVAL, "invalid racing game descriptor"); }
  if ((params2 ? dg_fill_params_v2(params2, &h->P) : dg_fill_params(params, &h->P)) != 0) { delete h; return set_err(DGSQP_EINVAL, "invalid solver parameters"); }
  h->D = make_dims(h->G.M, h->G.N);
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev) {
    delete h;
    return set_err(DGSQP_ECUDA, std::string("no usable CUDA device (") + cudaGetErrorString(e) +
                                    "); dgsqp_b200 has no CPU fallback");
  }
  h->device = device;
  int rc = 0;
  do {
    if (cudaSetDevice(device) != cudaSuccess) { rc = set_err(DGSQP_ECUDA, "cudaSetDevice failed"); break; }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { rc = set_err(DGSQP_ECUDA, "cudaGetDeviceProperties failed"); break; }
    h->sm_count = prop.multiProcessorCount;
    if (cudaMalloc(&h->d_G, sizeof(GameDesc)) != cudaSuccess || cudaMalloc(&h->d_P, sizeof(SolverParams)) != cudaSuccess ||
        cudaMalloc(&h->d_counter, sizeof(int)) != cudaSuccess) { rc = set_err(DGSQP_ENOMEM, "device allocation failed"); break; }
    cudaMemcpy(h->d_G, &h->G, sizeof(GameDesc), cudaMemcpyHostToDevice);
    cudaMemcpy(h->d_P, &h->P, sizeof(SolverParams), cudaMemcpyHostToDevice);
    // deep call chains with small local arrays
    cudaDeviceSetLimit(cudaLimitStackSize, 8192);
    rc = ensure_grid(h);
  } while (0);
  if (rc != 0) { dgsqp_destroy(h); return rc; }
  *out = h;
  return DGSQP_OK;
}

int dgsqp_create(const dgsqp_racing_game* game, const dgsqp_params* params, int device, dgsqp_handle** out) {
  return create_common(game, params, nullptr, device, out);
}

int dgsqp_create_v2(const dgsqp_racing_game* game, const dgsqp_v2_params* params, int device, dgsqp_handle** out) {
  if (!params) return set_err(DGSQP_EINVAL, "NULL parameters");
  return create_common(game, nullptr, params, device, out);
}

int dgsqp_destroy(dgsqp_handle* h) {
  if (!h) return DGSQP_OK;
  cudaSetDevice(h->device);
  cudaFree(h->d_ws); cudaFree(h->d_counter); cudaFree(h->d_diag); cudaFree(h->d_phase); cudaFree(h->d_G); cudaFree(h->d_P);
  free_stage(h);
  delete h;
  return DGSQP_OK;
}

int dgsqp_dims(const dgsqp_handle* h, int32_t dims[4]) {
  if (!h || !dims) return set_err(DGSQP_EINVAL, "NULL argument");
  dims[0] = h->D.nq; dims[1] = h->D.nu; dims[2] = h->D.n; dims[3] = h->D.m;
  return DGSQP_OK;
}

int dgsqp_configure(dgsqp_handle* h, int32_t ctas_per_sm, int32_t threads) {
  if (!h) return set_err(DGSQP_EINVAL, "NULL handle");
  if (threads != 0 && (threads < 32 || threads > DG_MAX_THREADS || (threads & 31))) return set_err(DGSQP_EINVAL, "threads must be a multiple of 32 in [32,256]");
  CUDA_TRY(cudaSetDevice(h->device));
  h->ctas_per_sm = ctas_per_sm;
  if (threads) h->threads = threads;
  return ensure_grid(h);
}

int dgsqp_solve_batch_async(dgsqp_handle* h, int32_t B, const double* x0, const double* u_ws, const double* l_ws, double* u_out,
                            double* l_out, double* x_out, double* cost_out, double* cond_out, int32_t* num_iters,
                            int32_t* status, int32_t* qp_solves, void* stream) {
  if (!h) return set_err(DGSQP_EINVAL, "NULL handle");
  if (B < 0) return set_err(DGSQP_EINVAL, "negative batch size");
  if (B == 0) return DGSQP_OK;
  if (!x0 || !u_ws || !u_out || !l_out || !x_out || !cost_out || !cond_out || !num_iters || !status || !qp_solves)
    return set_err(DGSQP_EINVAL, "NULL buffer");
  CUDA_TRY(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  if ((size_t)B > h->diag_cap) {
    if (h->d_diag) { cudaFree(h->d_diag); h->d_diag = nullptr; h->diag_cap = 0; }
    if (h->d_phase) { cudaFree(h->d_phase); h->d_phase = nullptr; }
    CUDA_TRY(cudaMalloc(&h->d_diag, sizeof(int) * DG_NDIAG * (size_t)B));
    CUDA_TRY(cudaMalloc(&h->d_phase, sizeof(long long) * DG_NPHASE * (size_t)B));
    h->diag_cap = (size_t)B;
  }
  CUDA_TRY(cudaMemsetAsync(h->d_counter, 0, sizeof(int), st));
  KernelArgs A;
  A.B = B; A.x0 = x0; A.u_ws = u_ws; A.l_ws = l_ws; A.u_out = u_out; A.l_out = l_out; A.x_out = x_out; A.cost_out = cost_out;
  A.cond_out = cond_out; A.num_iters = num_iters; A.status = status; A.qp_solves = qp_solves; A.diag = h->d_diag; A.phase = h->d_phase;
  A.ws = h->d_ws; A.ws_stride = h->ws_doubles; A.smem_doubles = h->smem_budget; A.counter = h->d_counter;
  A.smem_used = h->plan.smem;
  { const char* e = getenv("DGSQP_POISON"); A.poison = (e && e[0] == '1') ? 1 : 0; }
  int grid = B < h->grid_cap ? B : h->grid_cap;
  if (h->plan.hot_in_smem) dgsqp_solve_kernel<true><<<grid, h->threads, h->smem_bytes, st>>>(h->d_G, h->d_P, A);
  else dgsqp_solve_kernel<false><<<grid, h->threads, h->smem_bytes, st>>>(h->d_G, h->d_P, A);
  g_launches.fetch_add(1);
  CUDA_TRY(cudaGetLastError());
  return DGSQP_OK;
}

int dgsqp_solve_batch(dgsqp_handle* h, int32_t B, const double* x0, const double* u_ws, const double* l_ws, double* u_out, double* l_out,
                      double* x_out, double* cost_out, double* cond_out, int32_t* num_iters, int32_t* status,
                      int32_t* qp_solves, int32_t memspace, void* stream) {
  if (!h) return set_err(DGSQP_EINVAL, "NULL handle");
  if (memspace != 0 && memspace != 1) return set_err(DGSQP_EINVAL, "memspace must be 0 (host) or 1 (device)");
  cudaStream_t st = (cudaStream_t)stream;
  if (memspace == 1) {
    int rc = dgsqp_solve_batch_async(h, B, x0, u_ws, l_ws, u_out, l_out, x_out, cost_out, cond_out, num_iters, status, qp_solves, stream);
    if (rc != 0) return rc;
    CUDA_TRY(cudaStreamSynchronize(st));
    return DGSQP_OK;
  }
  if (B < 0) return set_err(DGSQP_EINVAL, "negative batch size");
  if (B == 0) return DGSQP_OK;
  if (!x0 || !u_ws || !u_out || !l_out || !x_out || !cost_out || !cond_out || !num_iters || !status || !qp_solves)
    return set_err(DGSQP_EINVAL, "NULL buffer");
  CUDA_TRY(cudaSetDevice(h->device));
  const Dims& D = h->D;
  if ((size_t)B > h->stage_cap) {
    free_stage(h);
    size_t b = (size_t)B;
    CUDA_TRY(cudaMalloc(&h->s_x0, sizeof(double) * b * D.nq));
    CUDA_TRY(cudaMalloc(&h->s_uws, sizeof(double) * b * D.n));
    CUDA_TRY(cudaMalloc(&h->s_lws, sizeof(double) * b * D.m));
    CUDA_TRY(cudaMalloc(&h->s_u, sizeof(double) * b * D.n));
    CUDA_TRY(cudaMalloc(&h->s_l, sizeof(double) * b * D.m));
    CUDA_TRY(cudaMalloc(&h->s_x, sizeof(double) * b * (D.N + 1) * D.nq));
    CUDA_TRY(cudaMalloc(&h->s_cost, sizeof(double) * b * D.M));
    CUDA_TRY(cudaMalloc(&h->s_cond, sizeof(double) * b * 3));
    CUDA_TRY(cudaMalloc(&h->s_it, sizeof(int) * b));
    CUDA_TRY(cudaMalloc(&h->s_st, sizeof(int) * b));
    CUDA_TRY(cudaMalloc(&h->s_qp, sizeof(int) * b));
    h->stage_cap = b;
  }
  size_t b = (size_t)B;
  CUDA_TRY(cudaMemcpyAsync(h->s_x0, x0, sizeof(double) * b * D.nq, cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaMemcpyAsync(h->s_uws, u_ws, sizeof(double) * b * D.n, cudaMemcpyHostToDevice, st));
  if (l_ws) CUDA_TRY(cudaMemcpyAsync(h->s_lws, l_ws, sizeof(double) * b * D.m, cudaMemcpyHostToDevice, st));
  int rc = dgsqp_solve_batch_async(h, B, h->s_x0, h->s_uws, l_ws ? h->s_lws : nullptr, h->s_u, h->s_l, h->s_x, h->s_cost, h->s_cond, h->s_it, h->s_st, h->s_qp, stream);
  if (rc != 0) return rc;
  CUDA_TRY(cudaMemcpyAsync(u_out, h->s_u, sizeof(double) * b * D.n, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaMemcpyAsync(l_out, h->s_l, sizeof(double) * b * D.m, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaMemcpyAsync(x_out, h->s_x, sizeof(double) * b * (D.N + 1) * D.nq, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaMemcpyAsync(cost_out, h->s_cost, sizeof(double) * b * D.M, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaMemcpyAsync(cond_out, h->s_cond, sizeof(double) * b * 3, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaMemcpyAsync(num_iters, h->s_it, sizeof(int) * b, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaMemcpyAsync(status, h->s_st, sizeof(int) * b, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaMemcpyAsync(qp_solves, h->s_qp, sizeof(int) * b, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  return DGSQP_OK;
}

int dgsqp_measure_fp64_peak(int device, double* tflops) {
  if (!tflops) return set_err(DGSQP_EINVAL, "NULL argument");
  CUDA_TRY(cudaSetDevice(device));
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, device));
  double* d_out = nullptr;
  CUDA_TRY(cudaMalloc(&d_out, sizeof(double)));
  cudaEvent_t e0, e1;
  CUDA_TRY(cudaEventCreate(&e0)); CUDA_TRY(cudaEventCreate(&e1));
  const int blocks = prop.multiProcessorCount * 4, threads = 512, iters = 1 << 16;
  double best = 0.0;
  for (int rep = 0; rep < 4; ++rep) {
    CUDA_TRY(cudaEventRecord(e0));
    dgsqp_fp64_probe_kernel<<<blocks, threads>>>(d_out, iters, 0.999999, 1e-9);
    g_launches.fetch_add(1);
    CUDA_TRY(cudaEventRecord(e1));
    CUDA_TRY(cudaEventSynchronize(e1));
    float ms = 0.f;
    CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
    double fl = 2.0 * 8.0 * (double)iters * (double)blocks * (double)threads;
    double tf = fl / (ms * 1e-3) / 1e12;
    if (rep > 0 && tf > best) best = tf;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d_out);
  *tflops = best;
  return DGSQP_OK;
}

int dgsqp_last_diag(dgsqp_handle* h, int32_t B, int32_t* diag) {
  if (!h || !diag) return set_err(DGSQP_EINVAL, "NULL argument");
  if (B < 0 || (size_t)B > h->diag_cap) return set_err(DGSQP_EINVAL, "B exceeds the last batch size");
  CUDA_TRY(cudaSetDevice(h->device));
  CUDA_TRY(cudaMemcpy(diag, h->d_diag, sizeof(int) * DG_NDIAG * (size_t)B, cudaMemcpyDeviceToHost));
  return DGSQP_OK;
}

int dgsqp_set_smem_limit(dgsqp_handle* h, int64_t bytes) {
  if (!h || bytes < 0) return set_err(DGSQP_EINVAL, "bad argument");
  CUDA_TRY(cudaSetDevice(h->device));
  h->smem_limit = bytes == 0 ? 0 : (size_t)bytes / sizeof(double) + 1;
  return ensure_grid(h);
}

int dgsqp_memory_plan(const dgsqp_handle* h, int64_t out[4]) {
  if (!h || !out) return set_err(DGSQP_EINVAL, "NULL argument");
  out[0] = (int64_t)(h->plan.smem * sizeof(double)); out[1] = (int64_t)(h->plan.gmem * sizeof(double));
  out[2] = h->plan.mats_in_smem; out[3] = h->plan.sens_in_smem + 2 * h->plan.hot_in_smem;
  return DGSQP_OK;
}

int dgsqp_phase_count(void) { return DG_NPHASE; }

int dgsqp_last_phase_cycles(dgsqp_handle* h, int32_t B, int64_t* cycles) {
  if (!h || !cycles) return set_err(DGSQP_EINVAL, "NULL argument");
  if (B < 0 || (size_t)B > h->diag_cap) return set_err(DGSQP_EINVAL, "B exceeds the last batch size");
  CUDA_TRY(cudaSetDevice(h->device));
  CUDA_TRY(cudaMemcpy(cycles, h->d_phase, sizeof(long long) * DG_NPHASE * (size_t)B, cudaMemcpyDeviceToHost));
  return DGSQP_OK;
}

}  // extern "C"
