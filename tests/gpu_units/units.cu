// Stage-level test kernel: runs evaluate -> nearestPD -> QP and the LSQR dual initialisation for a
// batch of instances (one CTA each) and returns every intermediate.  TEST ONLY (tests/test_gpu_stages.py).
#include <cuda_runtime.h>
#include <cstdio>
#include "../../dgsqp_b200/csrc/sqp_v1.cuh"
#include "../../dgsqp_b200/csrc/host_setup.h"

struct UnitArgs {
  int B; const double* x0; const double* u; const double* l;
  double *Q, *q, *gtl, *g, *H, *du, *lam, *l0;
  int *nneg, *qpst, *qpit, *lsqr_it;
  double* ws; size_t ws_stride; size_t smem_doubles;
};

template <bool SM>
__global__ void __launch_bounds__(256, 1) units_kernel(const GameDesc* Gp, const SolverParams* Pp, UnitArgs A) {
  extern __shared__ double s_dyn[];
  Cta c;
  c.flip = 0;
  __shared__ SolveCtx sX;
  if (threadIdx.x == 0) {
    sX.G = Gp; sX.P = Pp; sX.D = make_dims(Gp->M, Gp->N);
    plan_memory(sX.D, A.ws + (size_t)blockIdx.x * A.ws_stride, s_dyn, A.smem_doubles, sX.W);
  }
  __syncthreads();
  SolveCtx& X = sX;
  const Dims& D = X.D; const int n = D.n, m = D.m;
  game_row_table<SM>(c, D, X.W.E.rowtab);
  for (int inst = blockIdx.x; inst < A.B; inst += gridDim.x) {
    c.sync();
    if (threadIdx.x == 0) X.x0 = A.x0 + (size_t)inst * D.nq;
    const double* u = A.u + (size_t)inst * n; const double* l = A.l + (size_t)inst * m;
    DG_FOR(j, D.nu) X.W.S.up[j] = 0.0;
    c.sync();
    eval_full<SM>(c, X, u, l);
    c.sync();
    DG_FOR(t, n * n) A.Q[(size_t)inst * n * n + t] = X.W.E.Q[t];
    DG_FOR(t, n) { A.q[(size_t)inst * n + t] = X.W.E.q[t]; A.gtl[(size_t)inst * n + t] = X.W.E.gtl[t]; }
    DG_FOR(t, m) A.g[(size_t)inst * m + t] = X.W.E.g[t];
    int nneg = nearest_pd<SM>(c, n, X.W.E.Q, X.W.B, Pp->eig_floor, Pp->reg, Pp->conv_approx != 0);
    DG_FOR(t, n * n) A.H[(size_t)inst * n * n + t] = X.W.B.matA[(t / n) * D.ld + (t % n)];
    c.sync();
    int it = 0, na = 0;
    int st = qp_solve_gi<SM>(c, D, X.W.E, X.W.E.q, X.W.Q, X.W.B, &it, &na, 0);
    DG_FOR(t, n) A.du[(size_t)inst * n + t] = X.W.Q.xq[t];
    DG_FOR(t, m) A.lam[(size_t)inst * m + t] = X.W.Q.lam[t];
    c.sync();
    DG_FOR(t, m) X.W.S.l[t] = 0.0;
    c.sync();
    eval_grad<SM>(c, X, u, X.W.S.l, true);
    int li = lsqr_dual_init<SM>(c, D, X.W.E, X.W.L, X.W.E.q, X.W.S.l);
    DG_FOR(t, m) A.l0[(size_t)inst * m + t] = X.W.S.l[t];
    if (c.tid() == 0) { A.nneg[inst] = nneg; A.qpst[inst] = st; A.qpit[inst] = it; A.lsqr_it[inst] = li; }
    c.sync();
  }
}

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e)); return -2; } } while (0)

extern "C" int units_run(const dgsqp_racing_game* game, const dgsqp_params* params, int B, int threads, long smem_limit_doubles, const double* x0,
                         const double* u, const double* l, double* Q, double* q, double* gtl, double* g, double* H,
                         double* du, double* lam, double* l0, int* nneg, int* qpst, int* qpit, int* lsqr_it) {
  GameDesc G; SolverParams P;
  if (dg_fill_game(game, &G) || dg_fill_params(params, &P)) return -1;
  Dims D = make_dims(G.M, G.N);
  int optin = 0; CK(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, 0));
  cudaFuncAttributes fa; CK(cudaFuncGetAttributes(&fa, units_kernel<true>));
  size_t budget = ((size_t)optin - fa.sharedSizeBytes - 64) / sizeof(double);
  if (smem_limit_doubles > 0 && (size_t)smem_limit_doubles < budget) budget = (size_t)smem_limit_doubles;
  Workspace tmp; MemPlan pl = plan_memory(D, nullptr, nullptr, budget, tmp);
  size_t wsd = pl.gmem;
  const size_t n = D.n, m = D.m;
  GameDesc* dG; SolverParams* dP; double* ws;
  CK(cudaMalloc(&dG, sizeof(G))); CK(cudaMalloc(&dP, sizeof(P)));
  CK(cudaMemcpy(dG, &G, sizeof(G), cudaMemcpyHostToDevice)); CK(cudaMemcpy(dP, &P, sizeof(P), cudaMemcpyHostToDevice));
  int grid = B < 64 ? B : 64;
  CK(cudaMalloc(&ws, sizeof(double) * wsd * grid)); CK(cudaMemset(ws, 0, sizeof(double) * wsd * grid));
  UnitArgs A; A.B = B; A.ws = ws; A.ws_stride = wsd; A.smem_doubles = budget;
  double *dx0, *du_, *dl_;
  CK(cudaMalloc(&dx0, 8 * B * D.nq)); CK(cudaMalloc(&du_, 8 * B * n)); CK(cudaMalloc(&dl_, 8 * B * m));
  CK(cudaMemcpy(dx0, x0, 8 * B * D.nq, cudaMemcpyHostToDevice)); CK(cudaMemcpy(du_, u, 8 * B * n, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dl_, l, 8 * B * m, cudaMemcpyHostToDevice));
  A.x0 = dx0; A.u = du_; A.l = dl_;
  CK(cudaMalloc(&A.Q, 8 * B * n * n)); CK(cudaMalloc(&A.H, 8 * B * n * n)); CK(cudaMalloc(&A.q, 8 * B * n)); CK(cudaMalloc(&A.gtl, 8 * B * n));
  CK(cudaMalloc(&A.g, 8 * B * m)); CK(cudaMalloc(&A.du, 8 * B * n)); CK(cudaMalloc(&A.lam, 8 * B * m)); CK(cudaMalloc(&A.l0, 8 * B * m));
  CK(cudaMalloc(&A.nneg, 4 * B)); CK(cudaMalloc(&A.qpst, 4 * B)); CK(cudaMalloc(&A.qpit, 4 * B)); CK(cudaMalloc(&A.lsqr_it, 4 * B));
  cudaDeviceSetLimit(cudaLimitStackSize, 8192);
  size_t smem = sizeof(double) * pl.smem;
  if (pl.hot_in_smem) {
    CK(cudaFuncSetAttribute(units_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    units_kernel<true><<<grid, threads, smem>>>(dG, dP, A);
  } else {
    CK(cudaFuncSetAttribute(units_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    units_kernel<false><<<grid, threads, smem>>>(dG, dP, A);
  }
  CK(cudaGetLastError()); CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(Q, A.Q, 8 * B * n * n, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(H, A.H, 8 * B * n * n, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(q, A.q, 8 * B * n, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(gtl, A.gtl, 8 * B * n, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(g, A.g, 8 * B * m, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(du, A.du, 8 * B * n, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(lam, A.lam, 8 * B * m, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(l0, A.l0, 8 * B * m, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(nneg, A.nneg, 4 * B, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(qpst, A.qpst, 4 * B, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(qpit, A.qpit, 4 * B, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(lsqr_it, A.lsqr_it, 4 * B, cudaMemcpyDeviceToHost));
  cudaFree(dG); cudaFree(dP); cudaFree(ws); cudaFree(dx0); cudaFree(du_); cudaFree(dl_);
  cudaFree(A.Q); cudaFree(A.H); cudaFree(A.q); cudaFree(A.gtl); cudaFree(A.g); cudaFree(A.du); cudaFree(A.lam); cudaFree(A.l0);
  cudaFree(A.nneg); cudaFree(A.qpst); cudaFree(A.qpit); cudaFree(A.lsqr_it);
  return 0;
}
