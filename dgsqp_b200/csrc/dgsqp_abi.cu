// C-ABI entry points (include/dgsqp_b200.h).  This translation unit holds the engine of the racing games (kinematic
// bicycles on a curvature-segment track) and the game-independent entry points, which dispatch through the handle;
// dgsqp_merge_abi.cu holds the engine of the merge game.  See engine.inc.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <atomic>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <string>
#include <new>

#include "abi_common.h"

#define DGSQP_VERSION_STR "dgsqp_b200 0.2.0 (sm_100a)"

static thread_local std::string g_last_error;
std::atomic<long long> dg_launches{0};
int dg_set_err(int code, const std::string& msg) { g_last_error = msg; return code; }

namespace {
#include "sqp_v2.cuh"
#include "host_setup.h"
#include "engine.inc"
}  // namespace

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

// PID warm-start roll-outs: one thread per sampled agent, see pid_rollout.cuh
#include "pid_rollout.cuh"
__global__ void dgsqp_pid_rollout_kernel(RolloutParams P, int K, const double* __restrict__ s0, const double* __restrict__ xt0,
                                         const double* __restrict__ v0, double* __restrict__ q0, double* __restrict__ xy,
                                         double* __restrict__ u_ws) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= K) return;
  ro_agent(P, s0[i], xt0[i], v0[i], q0 + (size_t)i * 6, xy + (size_t)i * (P.N + 1) * 2, u_ws + (size_t)i * P.N * 2);
}

// Per-shard statistics of a solved batch, reduced on the device: the additive / max entries of the table
// scripts/process_data_curve.py:98-110 and process_data_merge.py:58-67 print.  out[16]: count, the six status counts,
// sum / sum of squares of SQP iterations, sum of QP solves, the same three over converged instances, max p_feas and max
// stat over converged instances, reserved.  One warp-shuffle + shared-memory reduction per block, one atomic per entry.
#define DG_NSTAT 16
__global__ void dgsqp_stats_kernel(const int* __restrict__ status, const int* __restrict__ iters, const int* __restrict__ qp,
                                   const double* __restrict__ cond, int B, double* __restrict__ out) {
  double v[DG_NSTAT];
  for (int k = 0; k < DG_NSTAT; ++k) v[k] = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < B; i += gridDim.x * blockDim.x) {
    const int st = status[i];
    const double it = (double)iters[i], q = (double)qp[i];
    v[0] += 1.0;
    if (st >= 0 && st < 6) v[1 + st] += 1.0;
    v[7] += it; v[8] += it * it; v[9] += q;
    if (st <= 1) {
      v[10] += it; v[11] += it * it; v[12] += q;
      if (cond) { v[13] = fmax(v[13], cond[3 * i]); v[14] = fmax(v[14], cond[3 * i + 2]); }
    }
  }
  __shared__ double sh[32][DG_NSTAT];
  for (int k = 0; k < DG_NSTAT; ++k) {
    double x = v[k];
    const bool is_max = k == 13 || k == 14;
    for (int o = 16; o > 0; o >>= 1) { const double y = __shfl_xor_sync(0xffffffffu, x, o); x = is_max ? fmax(x, y) : x + y; }
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5][k] = x;
  }
  __syncthreads();
  if (threadIdx.x < DG_NSTAT) {
    const int k = threadIdx.x;
    const bool is_max = k == 13 || k == 14;
    double x = 0.0;
    for (int w = 0; w < (int)((blockDim.x + 31) >> 5); ++w) x = is_max ? fmax(x, sh[w][k]) : x + sh[w][k];
    if (is_max) {
      // p_feas and |stat| are non-negative: the ordering of the bit patterns is the ordering of the values
      atomicMax((unsigned long long*)&out[k], (unsigned long long)__double_as_longlong(x));
    } else atomicAdd(&out[k], x);
  }
}

extern "C" {

int dgsqp_batch_stats(int device, int32_t B, const int32_t* status, const int32_t* num_iters, const int32_t* qp_solves,
                      const double* cond, double* out16, void* stream) {
  if (!status || !num_iters || !qp_solves || !out16) return dg_set_err(DGSQP_EINVAL, "NULL buffer");
  if (B < 0) return dg_set_err(DGSQP_EINVAL, "negative batch size");
  CUDA_TRY(cudaSetDevice(device));
  cudaStream_t st = (cudaStream_t)stream;
  double* d_out = nullptr;
  CUDA_TRY(cudaMalloc(&d_out, sizeof(double) * DG_NSTAT));
  cudaError_t e = cudaMemsetAsync(d_out, 0, sizeof(double) * DG_NSTAT, st);
  if (e == cudaSuccess && B > 0) {
    int blocks = (B + 255) / 256;
    if (blocks > 296) blocks = 296;
    dgsqp_stats_kernel<<<blocks, 256, 0, st>>>(status, num_iters, qp_solves, cond, B, d_out);
    dg_launches.fetch_add(1);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaMemcpyAsync(out16, d_out, sizeof(double) * DG_NSTAT, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  cudaFree(d_out);
  if (e != cudaSuccess) return dg_set_err(DGSQP_ECUDA, std::string("dgsqp_batch_stats: ") + cudaGetErrorString(e));
  return DGSQP_OK;
}

int dgsqp_pid_rollout(const dgsqp_racing_game* game, const double* key_pts, int device, int32_t K, const double* s0,
                      const double* xt0, const double* v0, double* q0, double* xy, double* u_ws, int32_t memspace, void* stream) {
  RolloutParams P;
  if (ro_fill(game, key_pts, &P) != 0) return dg_set_err(DGSQP_EINVAL, "invalid game / key points");
  if (K < 0 || (memspace != 0 && memspace != 1)) return dg_set_err(DGSQP_EINVAL, "bad argument");
  if (K == 0) return DGSQP_OK;
  if (!s0 || !xt0 || !v0 || !q0 || !xy || !u_ws) return dg_set_err(DGSQP_EINVAL, "NULL buffer");
  CUDA_TRY(cudaSetDevice(device));
  cudaStream_t st = (cudaStream_t)stream;
  const size_t k = (size_t)K, nxy = (size_t)(P.N + 1) * 2, nu = (size_t)P.N * 2;
  double* d = nullptr;          // host path: one staging block [s0 | xt0 | v0 | q0 | xy | u_ws]
  const double *ds0 = s0, *dxt = xt0, *dv0 = v0;
  double *dq0 = q0, *dxy = xy, *du = u_ws;
  if (memspace == 0) {
    CUDA_TRY(cudaMalloc(&d, sizeof(double) * k * (3 + 6 + nxy + nu)));
    double* in = d;
    dq0 = d + 3 * k; dxy = dq0 + 6 * k; du = dxy + nxy * k;
    cudaError_t e = cudaMemcpyAsync(in, s0, sizeof(double) * k, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(in + k, xt0, sizeof(double) * k, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(in + 2 * k, v0, sizeof(double) * k, cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) { cudaFree(d); return dg_set_err(DGSQP_ECUDA, std::string("dgsqp_pid_rollout: ") + cudaGetErrorString(e)); }
    ds0 = in; dxt = in + k; dv0 = in + 2 * k;
  }
  dgsqp_pid_rollout_kernel<<<(K + 127) / 128, 128, 0, st>>>(P, K, ds0, dxt, dv0, dq0, dxy, du);
  dg_launches.fetch_add(1);
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess && memspace == 0) {
    e = cudaMemcpyAsync(q0, dq0, sizeof(double) * k * 6, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(xy, dxy, sizeof(double) * k * nxy, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(u_ws, du, sizeof(double) * k * nu, cudaMemcpyDeviceToHost, st);
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (d) cudaFree(d);
  if (e != cudaSuccess) return dg_set_err(DGSQP_ECUDA, std::string("dgsqp_pid_rollout: ") + cudaGetErrorString(e));
  return DGSQP_OK;
}

const char* dgsqp_last_error(void) { return g_last_error.c_str(); }
const char* dgsqp_version(void) { return DGSQP_VERSION_STR; }
int64_t dgsqp_kernel_launches(void) { return (int64_t)dg_launches.load(); }

int dgsqp_create(const dgsqp_racing_game* game, const dgsqp_params* params, int device, dgsqp_handle** out) {
  return eng_create(game, params, nullptr, device, out);
}

int dgsqp_create_v2(const dgsqp_racing_game* game, const dgsqp_v2_params* params, int device, dgsqp_handle** out) {
  if (!params) return dg_set_err(DGSQP_EINVAL, "NULL parameters");
  return eng_create(game, nullptr, params, device, out);
}

int dgsqp_destroy(dgsqp_handle* h) { return h ? h->vt->destroy(h) : DGSQP_OK; }

int dgsqp_dims(const dgsqp_handle* h, int32_t dims[4]) {
  if (!h || !dims) return dg_set_err(DGSQP_EINVAL, "NULL argument");
  return h->vt->dims(h, dims);
}

int dgsqp_configure(dgsqp_handle* h, int32_t ctas_per_sm, int32_t threads) {
  if (!h) return dg_set_err(DGSQP_EINVAL, "NULL handle");
  return h->vt->configure(h, ctas_per_sm, threads);
}

int dgsqp_solve_batch_async(dgsqp_handle* h, int32_t B, const double* x0, const double* u_ws, const double* l_ws, double* u_out,
                            double* l_out, double* x_out, double* cost_out, double* cond_out, int32_t* num_iters,
                            int32_t* status, int32_t* qp_solves, void* stream) {
  if (!h) return dg_set_err(DGSQP_EINVAL, "NULL handle");
  return h->vt->solve_batch_async(h, B, x0, u_ws, l_ws, nullptr, u_out, l_out, x_out, cost_out, cond_out, num_iters, status, qp_solves, stream);
}

int dgsqp_solve_batch(dgsqp_handle* h, int32_t B, const double* x0, const double* u_ws, const double* l_ws, double* u_out, double* l_out,
                      double* x_out, double* cost_out, double* cond_out, int32_t* num_iters, int32_t* status,
                      int32_t* qp_solves, int32_t memspace, void* stream) {
  if (!h) return dg_set_err(DGSQP_EINVAL, "NULL handle");
  return h->vt->solve_batch(h, B, x0, u_ws, l_ws, nullptr, u_out, l_out, x_out, cost_out, cond_out, num_iters, status, qp_solves, memspace, stream);
}

int dgsqp_solve_batch_up(dgsqp_handle* h, int32_t B, const double* x0, const double* u_ws, const double* l_ws, const double* u_prev,
                         double* u_out, double* l_out, double* x_out, double* cost_out, double* cond_out, int32_t* num_iters,
                         int32_t* status, int32_t* qp_solves, int32_t memspace, void* stream) {
  if (!h) return dg_set_err(DGSQP_EINVAL, "NULL handle");
  return h->vt->solve_batch(h, B, x0, u_ws, l_ws, u_prev, u_out, l_out, x_out, cost_out, cond_out, num_iters, status, qp_solves, memspace, stream);
}

int dgsqp_measure_fp64_peak(int device, double* tflops) {
  if (!tflops) return dg_set_err(DGSQP_EINVAL, "NULL argument");
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
    dg_launches.fetch_add(1);
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
  if (!h || !diag) return dg_set_err(DGSQP_EINVAL, "NULL argument");
  return h->vt->last_diag(h, B, diag);
}

int dgsqp_set_smem_limit(dgsqp_handle* h, int64_t bytes) {
  if (!h || bytes < 0) return dg_set_err(DGSQP_EINVAL, "bad argument");
  return h->vt->set_smem_limit(h, bytes);
}

int dgsqp_memory_plan(const dgsqp_handle* h, int64_t out[4]) {
  if (!h || !out) return dg_set_err(DGSQP_EINVAL, "NULL argument");
  return h->vt->memory_plan(h, out);
}

int dgsqp_iter_log_capacity(const dgsqp_handle* h) {
  if (!h) return dg_set_err(DGSQP_EINVAL, "NULL handle");
  return h->vt->iter_log_capacity(h);
}

int dgsqp_last_iter_data(dgsqp_handle* h, int32_t B, double* out) {
  if (!h || !out) return dg_set_err(DGSQP_EINVAL, "NULL argument");
  return h->vt->last_iter_data(h, B, out);
}

int dgsqp_phase_count(void) { return DG_NPHASE; }

int dgsqp_last_stats(dgsqp_handle* h, double* out16) {
  if (!h) return dg_set_err(DGSQP_EINVAL, "NULL handle");
  return h->vt->last_stats(h, out16);
}

int dgsqp_last_phase_cycles(dgsqp_handle* h, int32_t B, int64_t* cycles) {
  if (!h || !cycles) return dg_set_err(DGSQP_EINVAL, "NULL argument");
  return h->vt->last_phase_cycles(h, B, cycles);
}

}  // extern "C"
