// Single-thread host build of the device solver source (DG_HOSTSIM).  TEST HARNESS ONLY: lets the
// CPU test-suite exercise the exact kernel source (arithmetic, control flow, memory bounds under
// ASan) on machines without a GPU.  It is not part of the product library and nothing in
// dgsqp_b200/ loads it.  Compiled once per game (-DDG_GAME_MERGE selects the merge game, see csrc/game.cuh).
#define DG_HOSTSIM 1
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../../dgsqp_b200/csrc/sqp_v2.cuh"
#include "../../dgsqp_b200/csrc/host_setup.h"

#ifndef DG_GAME_MERGE
#include "../../dgsqp_b200/csrc/pid_rollout.cuh"
extern "C" int hs_pid_rollout(const dgsqp_racing_game* g, const double* key_pts, int K, const double* s0, const double* xt0,
                              const double* v0, double* q0, double* xy, double* u_ws) {
  RolloutParams P;
  if (ro_fill(g, key_pts, &P) != 0) return -1;
  for (int i = 0; i < K; ++i) ro_agent(P, s0[i], xt0[i], v0[i], q0 + (size_t)i * 6, xy + (size_t)i * (P.N + 1) * 2, u_ws + (size_t)i * P.N * 2);
  return 0;
}
#endif

extern "C" {

struct HsHandle { GameDesc G; SolverParams P; Dims D; std::vector<double> ws, sh; Workspace W; std::vector<double*> guards; };
static const double HS_CANARY = -7.25e77;

static void* hs_create_common(const dg_game_struct* g, const dgsqp_params* p, const dgsqp_v2_params* p2) {
  HsHandle* h = new HsHandle();
  if (dg_fill_game(g, &h->G) != 0 || (p2 ? dg_fill_params_v2(p2, &h->P) : dg_fill_params(p, &h->P)) != 0) { delete h; return nullptr; }
  h->D = make_dims(h->G.M, h->G.N);
  Workspace tmp;
  const char* lim = getenv("DG_HOSTSIM_SMEM_DOUBLES");      // tests cover both placements
  size_t budget = lim ? (size_t)atol(lim) : 28000;
  MemPlan pl = plan_memory(h->D, nullptr, nullptr, budget, tmp);
  h->ws.assign(pl.gmem + 2, 0.0);
  h->sh.assign(pl.smem + 2, 0.0);
#ifdef DG_PLAN_GUARD
  dg_guard_n = 0;
#endif
  plan_memory(h->D, h->ws.data(), h->sh.data(), budget, h->W);
#ifdef DG_PLAN_GUARD
  // guard mode: every buffer of the plan is followed by DG_PLAN_GUARD canary doubles (see plan_memory)
  h->guards.assign(dg_guard_at, dg_guard_at + dg_guard_n);
  for (double* g : h->guards) for (int i = 0; i < DG_PLAN_GUARD; ++i) g[i] = HS_CANARY;
#endif
  game_bind(h->W.E, &h->G);
  { Cta c; game_row_table<false>(c, h->D, h->W.E.rowtab); }
  return h;
}
void* hs_create(const dg_game_struct* g, const dgsqp_params* p) { return hs_create_common(g, p, nullptr); }
void* hs_create_v2(const dg_game_struct* g, const dgsqp_v2_params* p) { return hs_create_common(g, nullptr, p); }
// number of canary doubles behind the buffers of the memory plan that no longer hold the canary (guard build only)
int hs_guard_check(void* hp) {
  HsHandle* h = (HsHandle*)hp;
  int bad = 0;
#ifdef DG_PLAN_GUARD
  for (double* g : h->guards) for (int i = 0; i < DG_PLAN_GUARD; ++i) bad += g[i] != HS_CANARY;
  if (h->guards.empty()) return -1;
#else
  bad = -1;
#endif
  return bad;
}
void hs_set_l0_perturb(void* hp, double v) { ((HsHandle*)hp)->P.dbg_l0_perturb = v; }
void hs_destroy(void* hp) { delete (HsHandle*)hp; }
void hs_dims(void* hp, int* out) { HsHandle* h = (HsHandle*)hp; out[0] = h->D.nq; out[1] = h->D.nu; out[2] = h->D.n; out[3] = h->D.m; }

// full evaluation at (u, l): returns Q[n*n], q[n], gtl[n], g[m], x
void hs_evaluate(void* hp, const double* x0, const double* u, const double* l, double* Q, double* q, double* gtl,
                 double* g, double* x) {
  HsHandle* h = (HsHandle*)hp; Cta c;
  SolveCtx X; X.G = &h->G; X.P = &h->P; X.D = h->D; X.W = h->W; X.x0 = x0; X.up_in = nullptr;
  for (int i = 0; i < h->D.nu; ++i) h->W.S.up[i] = 0.0;
  eval_full<false>(c, X, u, l);
  memcpy(Q, h->W.E.Q, sizeof(double) * h->D.n * h->D.n);
  memcpy(q, h->W.E.q, sizeof(double) * h->D.n);
  memcpy(gtl, h->W.E.gtl, sizeof(double) * h->D.n);
  memcpy(g, h->W.E.g, sizeof(double) * h->D.m);
  memcpy(x, h->W.E.x, sizeof(double) * (h->D.N + 1) * h->D.nq);
}
// dense G (m x n) at the last evaluated point, via row extraction; plus G v and G' w products
void hs_G_dense(void* hp, double* Gd) {
  HsHandle* h = (HsHandle*)hp; Cta c;
  for (int r = 0; r < h->D.m; ++r) game_G_row<false>(c, h->D, h->W.E, r, Gd + (size_t)r * h->D.n);
}
void hs_G_times(void* hp, const double* v, double* y) { HsHandle* h = (HsHandle*)hp; Cta c; game_G_times<false>(c, h->D, h->W.E, v, y); }
void hs_GT_times(void* hp, const double* w, double* y) { HsHandle* h = (HsHandle*)hp; Cta c; game_GT_times<false>(c, h->D, h->W.E, w, y); }

int hs_nearest_pd(void* hp, const double* Qin, double* Hout) {
  HsHandle* h = (HsHandle*)hp; Cta c;
  int nn = nearest_pd<false>(c, h->D.n, Qin, h->W.B, h->P.eig_floor, h->P.reg, h->P.conv_approx != 0);
  for (int i = 0; i < h->D.n; ++i) memcpy(Hout + (size_t)i * h->D.n, h->W.B.matA + (size_t)i * h->D.ld, sizeof(double) * h->D.n);
  return nn;
}
// QP at the last evaluated point with the given H (n*n, destroyed) and q
int hs_qp(void* hp, double* H, const double* q, double* du, double* lam, int* iters) {
  HsHandle* h = (HsHandle*)hp; Cta c;
  int na = 0;
  for (int i = 0; i < h->D.n; ++i) memcpy(h->W.B.matA + (size_t)i * h->D.ld, H + (size_t)i * h->D.n, sizeof(double) * h->D.n);
  std::vector<double> qc(q, q + h->D.n);      // q may alias workspace the solver reuses
  int st = qp_solve_gi<false>(c, h->D, h->W.E, qc.data(), h->W.Q, h->W.B, iters, &na, 0);
  memcpy(du, h->W.Q.xq, sizeof(double) * h->D.n);
  memcpy(lam, h->W.Q.lam, sizeof(double) * h->D.m);
  return st;
}
int hs_lsqr(void* hp, const double* x0, const double* u, double* l_out) {
  HsHandle* h = (HsHandle*)hp; Cta c;
  SolveCtx X; X.G = &h->G; X.P = &h->P; X.D = h->D; X.W = h->W; X.x0 = x0; X.up_in = nullptr;
  for (int i = 0; i < h->D.nu; ++i) h->W.S.up[i] = 0.0;
  std::vector<double> l0(h->D.m, 0.0);
  eval_grad<false>(c, X, u, l0.data(), true);
  return lsqr_dual_init<false>(c, h->D, h->W.E, h->W.L, h->W.E.q, l_out);
}
void hs_solve(void* hp, const double* x0, const double* u_ws, const double* l_ws, const double* u_prev, double* u, double* l, double* x, double* cost, double* cond,
              int* num_iters, int* status, int* qp_solves, int* diag, double* l_init) {
  HsHandle* h = (HsHandle*)hp; Cta c;
  SolveCtx X; X.G = &h->G; X.P = &h->P; X.D = h->D; X.W = h->W; X.x0 = x0; X.up_in = u_prev;
  SolveOut O; O.u = u; O.l = l; O.x = x; O.cost = cost; O.cond = cond; O.num_iters = num_iters; O.status = status;
  O.qp_solves = qp_solves; O.diag = diag; O.l_init = l_init; O.iter_log = nullptr; O.iter_cap = 0;
  if (h->P.policy == 2) sqp_solve_v2<false>(c, X, u_ws, l_ws, O);
  else sqp_solve_v1<false>(c, X, u_ws, l_ws, O);
}
}
