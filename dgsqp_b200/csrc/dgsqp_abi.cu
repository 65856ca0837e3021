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


extern "C" {

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
  return h->vt->solve_batch_async(h, B, x0, u_ws, l_ws, u_out, l_out, x_out, cost_out, cond_out, num_iters, status, qp_solves, stream);
}

int dgsqp_solve_batch(dgsqp_handle* h, int32_t B, const double* x0, const double* u_ws, const double* l_ws, double* u_out, double* l_out,
                      double* x_out, double* cost_out, double* cond_out, int32_t* num_iters, int32_t* status,
                      int32_t* qp_solves, int32_t memspace, void* stream) {
  if (!h) return dg_set_err(DGSQP_EINVAL, "NULL handle");
  return h->vt->solve_batch(h, B, x0, u_ws, l_ws, u_out, l_out, x_out, cost_out, cond_out, num_iters, status, qp_solves, memspace, stream);
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

int dgsqp_phase_count(void) { return DG_NPHASE; }

int dgsqp_last_phase_cycles(dgsqp_handle* h, int32_t B, int64_t* cycles) {
  if (!h || !cycles) return dg_set_err(DGSQP_EINVAL, "NULL argument");
  return h->vt->last_phase_cycles(h, B, cycles);
}

}  // extern "C"
