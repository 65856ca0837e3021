// Shared between the translation units of the library (dgsqp_abi.cu, dgsqp_merge_abi.cu): the handle's dispatch table,
// the error slot and the launch counter.
#pragma once
#include <atomic>
#include <string>
#include "../../include/dgsqp_b200.h"

struct dgsqp_handle_vtbl {
  int (*destroy)(dgsqp_handle*);
  int (*dims)(const dgsqp_handle*, int32_t*);
  int (*configure)(dgsqp_handle*, int32_t, int32_t);
  int (*solve_batch_async)(dgsqp_handle*, int32_t, const double*, const double*, const double*, const double*, double*, double*, double*, double*,
                           double*, int32_t*, int32_t*, int32_t*, void*);
  int (*solve_batch)(dgsqp_handle*, int32_t, const double*, const double*, const double*, const double*, double*, double*, double*, double*,
                     double*, int32_t*, int32_t*, int32_t*, int32_t, void*);
  int (*last_diag)(dgsqp_handle*, int32_t, int32_t*);
  int (*set_smem_limit)(dgsqp_handle*, int64_t);
  int (*memory_plan)(const dgsqp_handle*, int64_t*);
  int (*last_phase_cycles)(dgsqp_handle*, int32_t, int64_t*);
  int (*iter_log_capacity)(const dgsqp_handle*);
  int (*last_iter_data)(dgsqp_handle*, int32_t, double*);
  int (*last_stats)(dgsqp_handle*, double*);
};
struct dgsqp_handle { const dgsqp_handle_vtbl* vt = nullptr; };

int dg_set_err(int code, const std::string& msg);
extern std::atomic<long long> dg_launches;

#define CUDA_TRY(expr)                                                                           \
  do {                                                                                           \
    cudaError_t e__ = (expr);                                                                    \
    if (e__ != cudaSuccess)                                                                      \
      return dg_set_err(DGSQP_ECUDA, std::string(#expr) + ": " + cudaGetErrorString(e__));       \
  } while (0)
