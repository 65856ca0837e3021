// Latency / throughput micro-benchmarks for the FP64 solver design (one warp or one CTA, clock64 deltas).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/ubench/lat scripts/ubench/lat.cu
#include <cstdio>
#include <cuda_runtime.h>
#define REP 2048
__global__ void k_lat(double* out, long long* cyc, double a, double b, int nthreads_active) {
  __shared__ double sm[4096];
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = 1.0 + 1e-9 * i;
  __syncthreads();
  double x = a + threadIdx.x * 1e-12, y = b;
  long long t0, t1; int id = 0;
  // 0: dependent DFMA chain
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < REP; ++i) x = fma(x, y, a);
  t1 = clock64(); if (threadIdx.x == 0) cyc[id] = t1 - t0; id++;
  // 1: dependent DADD chain
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < REP; ++i) x = x + y;
  t1 = clock64(); if (threadIdx.x == 0) cyc[id] = t1 - t0; id++;
  // 2: dependent division chain
  t0 = clock64();
#pragma unroll 4
  for (int i = 0; i < REP; ++i) x = y / (x + 2.0);
  t1 = clock64(); if (threadIdx.x == 0) cyc[id] = t1 - t0; id++;
  // 3: dependent sqrt chain
  t0 = clock64();
#pragma unroll 4
  for (int i = 0; i < REP; ++i) x = sqrt(x + 2.0);
  t1 = clock64(); if (threadIdx.x == 0) cyc[id] = t1 - t0; id++;
  // 4: dependent reciprocal 1/x
  t0 = clock64();
#pragma unroll 4
  for (int i = 0; i < REP; ++i) x = 1.0 / (x + 2.0);
  t1 = clock64(); if (threadIdx.x == 0) cyc[id] = t1 - t0; id++;
  // 5: dependent shared load chain (pointer chasing through doubles)
  {
    int idx = threadIdx.x & 31;
    t0 = clock64();
#pragma unroll 8
    for (int i = 0; i < REP; ++i) idx = (int)(sm[idx & 4095] * 0.0) + ((idx + 33) & 4095);
    t1 = clock64(); if (threadIdx.x == 0) cyc[id] = t1 - t0; id++;
    x += idx;
  }
  // 6: double shuffle-xor chain (one butterfly step each)
  t0 = clock64();
#pragma unroll 8
  for (int i = 0; i < REP; ++i) x += __shfl_xor_sync(0xffffffffu, x, 1 << (i % 5));
  t1 = clock64(); if (threadIdx.x == 0) cyc[id] = t1 - t0; id++;
  // 7: __syncthreads back to back
  t0 = clock64();
  for (int i = 0; i < REP; ++i) __syncthreads();
  t1 = clock64(); if (threadIdx.x == 0) cyc[id] = t1 - t0; id++;
  // 8: 8 independent DFMA chains (throughput per warp)
  {
    double z0 = x, z1 = x + 1, z2 = x + 2, z3 = x + 3, z4 = x + 4, z5 = x + 5, z6 = x + 6, z7 = x + 7;
    t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < REP; ++i) { z0 = fma(z0, y, a); z1 = fma(z1, y, a); z2 = fma(z2, y, a); z3 = fma(z3, y, a); z4 = fma(z4, y, a); z5 = fma(z5, y, a); z6 = fma(z6, y, a); z7 = fma(z7, y, a); }
    t1 = clock64(); if (threadIdx.x == 0) cyc[id] = t1 - t0; id++;
    x += z0 + z1 + z2 + z3 + z4 + z5 + z6 + z7;
  }
  // 9: sincos chain
  t0 = clock64();
  for (int i = 0; i < REP / 8; ++i) { double s, c; sincos(x, &s, &c); x = s + c; }
  t1 = clock64(); if (threadIdx.x == 0) cyc[id] = t1 - t0; id++;
  // 10: fmod chain
  t0 = clock64();
  for (int i = 0; i < REP / 8; ++i) x = fmod(x + 17.3, 15.0);
  t1 = clock64(); if (threadIdx.x == 0) cyc[id] = t1 - t0; id++;
  // 11: atan2 chain
  t0 = clock64();
  for (int i = 0; i < REP / 8; ++i) x = atan2(x, 0.26);
  t1 = clock64(); if (threadIdx.x == 0) cyc[id] = t1 - t0; id++;
  // 12: shared load + FMA stream (independent): sum += sm[i*33 + lane] * y
  {
    double acc0 = 0, acc1 = 0;
    t0 = clock64();
#pragma unroll 8
    for (int i = 0; i < 64; ++i) { acc0 += sm[(i * 64 + (threadIdx.x & 31)) & 4095] * y; acc1 += sm[(i * 64 + 32 + (threadIdx.x & 31)) & 4095] * y; }
    t1 = clock64(); if (threadIdx.x == 0) cyc[id] = t1 - t0; id++;
    x += acc0 + acc1;
  }
  out[threadIdx.x] = x;
}
int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 1024 * 8); cudaMalloc(&cyc, 64 * 8);
  const char* names[] = {"DFMA dep", "DADD dep", "DDIV dep (+1 add)", "DSQRT dep (+1 add)", "DRCP dep (+1 add)", "LDS dep (+cvt,+int ops)", "SHFL.f64 dep (+add)",
                         "__syncthreads", "DFMA x8 indep (per 8)", "sincos dep", "fmod dep", "atan2 dep", "LDS+DFMA stream (per 128 elems)"};
  const int reps[] = {REP, REP, REP, REP, REP, REP, REP, REP, REP, REP / 8, REP / 8, REP / 8, 1};
  for (int nt : {32, 256, 512}) {
    k_lat<<<1, nt>>>(out, cyc, 1.0000001, 0.9999999, nt);
    cudaDeviceSynchronize();
    long long h[16]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    printf("threads %d\n", nt);
    for (int i = 0; i < 13; ++i) printf("  %-34s %8.1f cycles/op\n", names[i], (double)h[i] / reps[i]);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
