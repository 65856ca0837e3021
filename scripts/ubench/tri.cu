// Timing harness for sym_tridiag_regs / sym_tridiag (one CTA, clock64), with knobs that switch parts of a step off.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -DDG_TRI_KNOBS -o scripts/ubench/tri scripts/ubench/tri.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
__device__ int dg_tri_knob = 0;
#include "../../dgsqp_b200/csrc/linalg.cuh"

__global__ void __launch_bounds__(256, 1) k_tri(int n, int ld, const double* A, double* out, long long* cyc, int variant, int reps) {
  extern __shared__ double sm[];
  Cta c; c.flip = 0;
  LinBuf B;
  B.ld = ld; B.matA = sm; B.matB = sm + n * ld;
  double* z = sm + 2 * n * ld;
  B.dg = z; B.od = z + n; B.od2 = z + 2 * n; B.tau = z + 3 * n; B.lam = z + 4 * n; B.pv = z + 5 * n; B.wv = z + 6 * n; B.sp = z;
  B.part = z + 8 * n; B.Zg = nullptr;
  long long tot = 0;
  for (int rep = 0; rep < reps; ++rep) {
    for (int t = threadIdx.x; t < n * n; t += blockDim.x) { int i = t / n, j = t - i * n; sm[i * ld + j] = A[t]; }
    __syncthreads();
    long long t0 = clock64();
    if (variant == 0) sym_tridiag_regs<52, true>(c, n, B);
    else sym_tridiag<true>(c, n, B);
    long long t1 = clock64();
    tot += t1 - t0;
    __syncthreads();
  }
  if (threadIdx.x == 0) cyc[0] = tot / reps;
  for (int t = threadIdx.x; t < n; t += blockDim.x) { out[t] = B.dg[t]; out[n + t] = B.od[t]; }
}

int main(int argc, char** argv) {
  const int n = argc > 1 ? atoi(argv[1]) : 100, ld = n | 1;
  double* hA = (double*)malloc(sizeof(double) * n * n);
  srand(1);
  for (int i = 0; i < n; ++i) for (int j = 0; j <= i; ++j) { double v = (double)rand() / RAND_MAX - 0.5; hA[i * n + j] = v; hA[j * n + i] = v; }
  double *dA, *dout; long long* dc;
  cudaMalloc(&dA, sizeof(double) * n * n); cudaMalloc(&dout, sizeof(double) * 2 * n); cudaMalloc(&dc, 8);
  cudaMemcpy(dA, hA, sizeof(double) * n * n, cudaMemcpyHostToDevice);
  size_t smem = sizeof(double) * (2 * n * ld + 8 * n + 512 + 64);
  cudaFuncSetAttribute(k_tri, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  double ref[2][256];
  for (int variant = 0; variant < 2; ++variant) {
    for (int knob = 0; knob < 1; ++knob) {
      cudaMemcpyToSymbol(dg_tri_knob, &knob, sizeof(int));
      k_tri<<<1, 256, smem>>>(n, ld, dA, dout, dc, variant, 20);
      cudaError_t e = cudaDeviceSynchronize();
      long long cyc; cudaMemcpy(&cyc, dc, 8, cudaMemcpyDeviceToHost);
      double h[256]; cudaMemcpy(h, dout, sizeof(double) * 2 * n, cudaMemcpyDeviceToHost);
      if (knob == 0) for (int i = 0; i < 2 * n; ++i) ref[variant][i] = h[i];
      printf("variant %d (%s) knob %2d: %8lld cycles per tridiagonalisation (%.0f per step)  dg[0..2] %.6f %.6f %.6f  od[0..1] %.6f %.6f  %s\n", variant,
             variant == 0 ? "regs" : "smem", knob, cyc, (double)cyc / (n - 1), h[0], h[1], h[2], h[n], h[n + 1], cudaGetErrorString(e));
    }
  }
  double md = 0; for (int i = 0; i < 2 * n; ++i) { double d = fabs(fabs(ref[0][i]) - fabs(ref[1][i])); if (d > md) md = d; }
  printf("max | |regs| - |smem| | over dg, od: %.3e\n", md);
  return 0;
}
