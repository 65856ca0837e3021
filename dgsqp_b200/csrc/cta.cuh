// CTA-cooperative execution context.
//
// Every solver routine in this directory is written as a function executed by ALL threads of
// one thread block (one game instance per CTA): strided loops over `c.tid`, `c.sync()` between
// dependent phases, block-wide reductions through `c.sum/max/argmin`.  Control flow is uniform:
// every branch is decided from values all threads read identically (shared memory or reduction
// results).
//
// The same source also compiles with a plain C++ compiler (DG_HOSTSIM) as a one-thread "CTA";
// that build exists only for tests/hostsim (algorithm debugging under ASan on machines without a
// GPU).  It is never linked into the product library.
#pragma once
#include <math.h>
#include <stdint.h>

#ifdef DG_HOSTSIM
#define DG_DEV static inline
#define DG_HD static inline
#define DG_DEVN static
#define DG_CONST static const
#define DG_RESTRICT
struct Cta {
  int tid = 0, nt = 1, lane = 0, warp = 0, nwarps = 1;
  static constexpr int wsz = 1;   // lanes per warp
  double* red = nullptr;
  inline void sync() {}
  inline double sum(double v) { return v; }
  inline double max(double v) { return v; }
  inline double min(double v) { return v; }
  inline int imin(int v) { return v; }
  inline int isum(int v) { return v; }
  // smallest value, ties -> smallest index; every thread gets the winner
  inline void argmin(double v, int idx, double& ov, int& oi) { ov = v; oi = idx; }
  inline double warp_sum(double v) { return v; }
};
#else
#define DG_DEV __device__ __forceinline__
#define DG_HD __host__ __device__ __forceinline__
#define DG_DEVN __device__ __noinline__
#define DG_CONST __device__ const
#define DG_RESTRICT __restrict__
struct Cta {
  int tid, nt, lane, warp, nwarps;
  static constexpr int wsz = 32;  // lanes per warp
  double* red;   // shared scratch, >= 2*nwarps+2 doubles
  __device__ __forceinline__ void sync() { __syncthreads(); }
  __device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
  }
  __device__ __forceinline__ double sum(double v) {
    v = warp_sum(v);
    __syncthreads();                       // protect `red` from the previous reduction's readers
    if (lane == 0) red[warp] = v;
    __syncthreads();
    double r = 0.0;
    for (int w = 0; w < nwarps; ++w) r += red[w];   // same order in every thread -> identical result
    return r;
  }
  __device__ __forceinline__ double max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    double r = red[0];
    for (int w = 1; w < nwarps; ++w) r = fmax(r, red[w]);
    return r;
  }
  __device__ __forceinline__ double min(double v) { return -max(-v); }
  __device__ __forceinline__ int imin(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { int t = __shfl_xor_sync(0xffffffffu, v, o); v = t < v ? t : v; }
    __syncthreads();
    if (lane == 0) ((int*)red)[warp] = v;
    __syncthreads();
    int r = ((int*)red)[0];
    for (int w = 1; w < nwarps; ++w) { int t = ((int*)red)[w]; r = t < r ? t : r; }
    return r;
  }
  __device__ __forceinline__ int isum(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if (lane == 0) ((int*)red)[warp] = v;
    __syncthreads();
    int r = 0;
    for (int w = 0; w < nwarps; ++w) r += ((int*)red)[w];
    return r;
  }
  __device__ __forceinline__ void argmin(double v, int idx, double& ov, int& oi) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      double v2 = __shfl_xor_sync(0xffffffffu, v, o);
      int i2 = __shfl_xor_sync(0xffffffffu, idx, o);
      if (v2 < v || (v2 == v && i2 < idx)) { v = v2; idx = i2; }
    }
    __syncthreads();
    if (lane == 0) { red[warp] = v; ((int*)(red + nwarps))[warp] = idx; }
    __syncthreads();
    ov = red[0]; oi = ((int*)(red + nwarps))[0];
    for (int w = 1; w < nwarps; ++w) {
      double v2 = red[w]; int i2 = ((int*)(red + nwarps))[w];
      if (v2 < ov || (v2 == ov && i2 < oi)) { ov = v2; oi = i2; }
    }
  }
};
#endif

#define DG_FOR(i, n) for (int i = c.tid; i < (n); i += c.nt)
