// DG-SQP v1 outer loop for ONE game instance, executed by one CTA.
//
// Restates DGSQP.solve (DGSQP/solvers/DGSQP.py:302-507) with its helpers _get_mu (:559-585), the
// merit functions f_phi / f_dphi / f_dstat_norm (:962-979), _line_search_3 (:1057-1081) and
// _watchdog_line_search_4 (:1174-1288).  Everything -- rollout, KKT assembly, eigen-projection,
// QP, merit, step acceptance, convergence tests -- runs on device; the host only launches.
#pragma once
#include "game.cuh"
#include "linalg.cuh"
#include "qp_gi.cuh"
#include "lsqr.cuh"

#define DG_NDIAG 8
enum { ST_CONV_ABS = 0, ST_CONV_REL = 1, ST_MAX_IT = 2, ST_DIVERGED = 3, ST_QP_FAIL = 4, ST_TIME_LIMIT = 5 };

struct SolverParams {
  double reg, p_tol, d_tol, beta, tau, eig_floor, merit_max, diverge_tol;
  double mu_vio_thresh;    // _get_mu's `thresh` (DGSQP.py:560); see DESIGN.md deviation D2
  double dbg_l0_perturb;   // test hook (sensitivity studies): relative perturbation of the dual initialisation; 0 in production
  int line_search_iters, sqp_iters, nonmono_ls, merit_l1, conv_approx, rel_tol_req, t_hat;
  // v2 step policy (DGSQPV2Params, DGSQP_v2.py:55-230); policy = 1: v1 (sqp_v1.cuh), 2: v2 (sqp_v2.cuh)
  int policy, nms, nms_frequency, nms_memory, armijo, has_merit_parameter;
  int merit_obj;           // v2 merit 'sum_obj_l1' (sum of the agents' costs) instead of 'stat_l1'
  int qp_warm;             // warm-start the active-set QP from the active set of the instance's previous QP (qp_gi.cuh: gi_warm_start)
  int iter_log;            // keep (p_feas, comp, stat, qp_solves) per SQP iteration (SolveOut::iter_log)
  double time_limit_ns;    // DGSQPParams.time_limit (DGSQP.py:470-474) per instance in ns of %globaltimer; 0 = none
  double reg_decay, sigma, gamma, merit_parameter;
};

struct SqpBuf {
  double *u, *l, *u_im1, *l_im1;
  double *du, *dl, *s, *ds;          // step of the outer iteration (kept for the watchdog fallback)
  double *u_t, *l_t, *du_t, *dl_t, *s_t, *ds_t;
  double *u_c, *l_c;                 // candidate point
  double *Gdu, *tn, *tn2;
  double* up;                        // nu (zeros: v1 resets u_prev every solve, DGSQP.py:305)
  // v2 only: iterate at the start of the running iteration, record of the last appended iteration
  double *c_u, *c_l, *r_u, *r_du, *r_l, *r_dl, *r_s, *r_ds;
  double* qs;                        // n: grad_u sum_f J^f at the evaluated point (v2 merit 'sum_obj_l1' only)
};

struct Workspace { EvalBuf E; LinBuf B; QpBuf Q; LsqrBuf L; SqpBuf S; };

#ifdef DG_PLAN_GUARD
static double* dg_guard_at[512];
static int dg_guard_n = 0;
#endif

struct MemPlan { size_t smem, gmem; int mats_in_smem, matA_in_smem, sens_in_smem, hot_in_smem; };   // doubles used in each space

// Memory plan of one CTA (= one game instance in flight).  Everything a latency-bound phase walks through
// lives in shared memory when it fits (budget `sbudget` doubles), the rest in the CTA's slice of global
// workspace.  Layout decisions, in priority order:
//   POOL   small hot vectors: trajectory, constraint values, gradients, the phase-aliased 8n region
//          (tridiagonal data | Cholesky panel | active-set vectors), slacks, active-set bookkeeping
//   ARENA  two n x ld matrices matA | matB (see LinBuf).  While the game is being evaluated no matrix is alive
//          and the arena holds the per-stage derivative data instead (AB, costates, contracted second
//          derivatives, value-function Hessians, the Hessian-DP rows)
//   SENS   packed sensitivity rows (alive from the evaluation through the QP and the merit computation)
//   then the SQP iterate vectors while room remains.
// The raw game Hessian Q, the LSQR Krylov basis and the cold SQP vectors stay in global memory (streamed
// with coalesced accesses, never on a dependent chain).
// Called with null bases to measure; decisions depend only on (D, sbudget).
DG_HD MemPlan plan_memory(const Dims& D, double* gbase, double* sbase, size_t sbudget, Workspace& W) {
  size_t so = 0, go = 0;
  const size_t n = D.n, m = D.m, nq = D.nq, N = D.N, M = D.M, ld = D.ld;
  auto rnd = [](size_t c) { return (c + 1) & ~(size_t)1; };
  // DG_PLAN_GUARD (tests/hostsim only): every buffer is followed by a gap of DG_PLAN_GUARD doubles whose position is
  // recorded, so that the test harness can fill the gaps with a canary and detect writes past the end of a buffer
#ifdef DG_PLAN_GUARD
#define DG_GPAD ((size_t)(DG_PLAN_GUARD))
#define DG_GNOTE(p, cnt) do { if ((p) && dg_guard_n < 512) dg_guard_at[dg_guard_n++] = (p) + rnd(cnt); } while (0)
#else
#define DG_GPAD ((size_t)0)
#define DG_GNOTE(p, cnt) do { } while (0)
#endif
#define GTAKE(ptr, cnt) do { (ptr) = gbase ? gbase + go : nullptr; go += rnd(cnt) + DG_GPAD; DG_GNOTE(ptr, cnt); } while (0)
#define STAKE(ptr, cnt) do { (ptr) = sbase ? sbase + so : nullptr; so += rnd(cnt) + DG_GPAD; DG_GNOTE(ptr, cnt); } while (0)
#define PLACE(ptr, cnt) do { if (so + rnd(cnt) + DG_GPAD <= sbudget) STAKE(ptr, cnt); else GTAKE(ptr, cnt); } while (0)
  MemPlan P; P.mats_in_smem = 0; P.sens_in_smem = 0;
  // ---- POOL
  PLACE(W.E.x, (N + 1) * nq); PLACE(W.E.g, m); PLACE(W.E.q, n); PLACE(W.E.gtl, n);
  PLACE(W.E.tmpS, M * N * 3); PLACE(W.E.cf, M * N * 3); PLACE(W.S.up, D.nu);
  {
    double* reg8 = nullptr;
    PLACE(reg8, 8 * n);
    double* z = reg8;
    auto at = [&](size_t k) { return z ? z + k * n : nullptr; };
    W.B.dg = at(0); W.B.od = at(1); W.B.od2 = at(2); W.B.tau = at(3); W.B.lam = at(4); W.B.pv = at(5); W.B.wv = at(6);
    W.B.sp = at(0);                                              // n * DG_CHOL_NB, DG_CHOL_NB == 8
    W.Q.xq = at(0); W.Q.dv = at(1); W.Q.zv = at(2); W.Q.rv = at(3); W.Q.npv = at(4); W.Q.lam_act = at(5);
  }
  PLACE(W.Q.sl, m);
  PLACE(W.B.part, (2 * n > DG_PART_SZ ? 2 * n : DG_PART_SZ));     // also holds the 2(n-1) Givens parameters of a drop step
  { double* t; PLACE(t, (n + 1) / 2 + 1); W.Q.act = (int*)t; PLACE(t, (m + 1) / 2 + 1); W.Q.is_act = (int*)t; }
  const bool pool_ok = go == 0;                 // nothing of the pool fell back to global memory
  // ---- ARENA: matA | matB.  Three placements: both matrices in shared memory; matA alone (split: matA carries the
  // latency-critical work -- reflectors, Cholesky factor, the triangular factor R of the active-set loop -- while matB
  // (eigenvector scratch, J' of the QP) is streamed with coalesced CTA-wide sweeps from the L2-resident workspace);
  // or both in the global workspace.
  const size_t ev_sz[7] = {rnd(N * M * DG_AB_SZ), rnd((M + 1) * (N + 1) * nq), rnd((M + 1) * N * M * DG_HC_SZ),
                           rnd(2 * (M + 1) * nq * nq + (M + 1) * D.nu * nq), rnd((M + 1) * nq * n), rnd(N * M * DG_T2_SZ), rnd(m)};
  size_t ev_a = 0;
  for (int i = 0; i < 7; ++i) ev_a += ev_sz[i];
  // matB doubles as the scratch of the eigenvector stage of nearest_pd (Sturm counters, DG_EIG_CHUNK interleaved
  // inverse-iteration work vectors and iterates: (n+1)/2+1 + 6 n DG_EIG_CHUNK doubles), which exceeds n*ld below n ~ 97
  const size_t eig_scr = rnd((n + 1) / 2 + 1 + 6 * n * DG_EIG_CHUNK);
  const size_t szA = rnd(n * ld);
  size_t szB = rnd(n * ld) > eig_scr ? rnd(n * ld) : eig_scr;
  if (szA + szB < ev_a) szB = ev_a - szA;
  double *arA = nullptr, *arB = nullptr;
  double** ev_ptr[7] = {&W.E.AB, &W.E.cst, &W.E.Hc, &W.E.Vbuf, &W.E.Wrow, &W.E.T2, &W.E.lbuf};
  P.matA_in_smem = 0;
  // (the split is only taken when the sensitivity rows still fit beside matA: every active-set iteration walks them)
  if (so + szA + szB + DG_GPAD <= sbudget || so + szA + rnd(M * D.sens_sz) + 2 * DG_GPAD > sbudget) {
    // one block: the evaluation data is carved from its start (no matrix is alive while the game is evaluated)
    if (so + szA + szB + DG_GPAD <= sbudget) { STAKE(arA, szA + szB); P.mats_in_smem = 1; P.matA_in_smem = 1; } else GTAKE(arA, szA + szB);
    arB = arA ? arA + szA : nullptr;
    size_t o = 0;
    for (int i = 0; i < 7; ++i) { *ev_ptr[i] = arA ? arA + o : nullptr; o += ev_sz[i]; }
  } else {
    // split: the evaluation buffers may not straddle the two spaces; small serial-chain data first into matA's block
    const int order[7] = {0, 1, 2, 3, 6, 5, 4};
    size_t oa = 0, ob = 0;
    for (int t = 0; t < 7; ++t) { const int i = order[t]; if (oa + ev_sz[i] <= szA) oa += ev_sz[i]; else ob += ev_sz[i]; }
    if (ob > szB) szB = ob;
    STAKE(arA, szA); GTAKE(arB, szB); P.matA_in_smem = 1;
    oa = ob = 0;
    for (int t = 0; t < 7; ++t) {
      const int i = order[t];
      if (oa + ev_sz[i] <= szA) { *ev_ptr[i] = arA ? arA + oa : nullptr; oa += ev_sz[i]; }
      else { *ev_ptr[i] = arB ? arB + ob : nullptr; ob += ev_sz[i]; }
    }
  }
  W.B.ld = (int)ld; W.B.matA = arA; W.B.matB = arB;
  // ---- SENS
  if (so + rnd(M * D.sens_sz) + DG_GPAD <= sbudget) { STAKE(W.E.S, M * D.sens_sz); P.sens_in_smem = 1; } else GTAKE(W.E.S, M * D.sens_sz);
  // ---- SQP iterate vectors, hottest first
  PLACE(W.S.u, n); PLACE(W.S.du, n); PLACE(W.S.l, m); PLACE(W.S.dl, m); PLACE(W.Q.lam, m);
  PLACE(W.S.u_c, n); PLACE(W.S.l_c, m); PLACE(W.S.s, m); PLACE(W.S.ds, m); PLACE(W.S.Gdu, m);
  PLACE(W.S.tn, n); PLACE(W.S.tn2, n);
  // ---- small shared-memory scratch of the inverse iteration when matB is not shared-memory resident (linalg.cuh: eig_s)
  W.B.eig_s = nullptr;
  if (!P.mats_in_smem && so + rnd(6 * n * DG_EIG_SMALL) + DG_GPAD <= sbudget) STAKE(W.B.eig_s, 6 * n * DG_EIG_SMALL);
  // ---- global only
  GTAKE(W.E.Q, n * n); GTAKE(W.B.Zg, n * n);
  { double* t; GTAKE(t, (m + 1) / 2 + 1); W.E.rowtab = (int*)t; }
  GTAKE(W.L.Ub, DG_LSQR_BASIS * m); GTAKE(W.L.Vb, DG_LSQR_BASIS * m); GTAKE(W.L.cf, DG_LSQR_BASIS);
  GTAKE(W.L.u, m); GTAKE(W.L.v, m); GTAKE(W.L.w, m); GTAKE(W.L.x, m); GTAKE(W.L.tn, n); GTAKE(W.L.tm, m);
  GTAKE(W.S.u_im1, n); GTAKE(W.S.l_im1, m);
  GTAKE(W.S.u_t, n); GTAKE(W.S.l_t, m); GTAKE(W.S.du_t, n); GTAKE(W.S.dl_t, m); GTAKE(W.S.s_t, m); GTAKE(W.S.ds_t, m);
  GTAKE(W.S.c_u, n); GTAKE(W.S.c_l, m); GTAKE(W.S.r_u, n); GTAKE(W.S.r_du, n); GTAKE(W.S.r_l, m); GTAKE(W.S.r_dl, m);
  GTAKE(W.S.r_s, m); GTAKE(W.S.r_ds, m);
  GTAKE(W.S.qs, n);
#undef GTAKE
#undef DG_GPAD
#undef DG_GNOTE
#undef STAKE
#undef PLACE
  P.smem = so; P.gmem = go;
  P.hot_in_smem = pool_ok && P.mats_in_smem && P.sens_in_smem;
  return P;
}

struct SolveCtx {
  const GameDesc* G;
  const SolverParams* P;
  Dims D;
  Workspace W;
  const double* x0;
  const double* up_in;   // previous input u_{-1} of this instance [nu] or null (v2 only: DGSQP_v2.py:328 keeps u_prev; v1 zeroes it, DGSQP.py:305)
  // work counters (maintained by thread 0: the context is shared by the CTA)
  int n_evals_full, n_evals_grad, n_gi_iters, n_neg_max, n_qp_indef, n_neg_sum, n_act_sum, n_ls_trials;
};

// _evaluate(u, l, hessian=True)
template <bool SM>
DG_DEVN void eval_full(Cta& c, SolveCtx& X, const double* u, const double* l_in) {
  const Dims D = X.D; const EvalBuf E = X.W.E; DG_SH_EVAL(E);
  c.sync();
  c.lap(PH_OTHER);
  // the multipliers are walked by serial chains (costates, Hessian DP): stage them in the arena (the candidate /
  // watchdog iterates live in global memory)
  { double* lb = E.lbuf; DG_FOR(r, D.m) lb[r] = l_in[r]; }
  const double* l = E.lbuf;
  game_rollout<SM>(c, *X.G, D, u, X.x0, E.x, E.tmpS);
  c.sync();
  game_linearize<SM>(c, *X.G, D, u, E, true);
  c.sync();
  c.lap(PH_LIN_FULL);
  game_constraints<SM>(c, *X.G, D, u, X.W.S.up, E.x, E.g, E.rowtab);
  game_costates<SM>(c, *X.G, D, E, l);
  game_sens<SM>(c, D, E);
  c.sync();
  game_contract<SM>(c, D, E);
  game_gradients<SM>(c, *X.G, D, E, u, X.W.S.up, l, X.P->merit_obj ? X.W.S.qs : nullptr);
  c.sync();
  c.lap(PH_ADJ_FULL);
  game_hessian<SM>(c, *X.G, D, E, l);
  c.lap(PH_HESS);
  if (c.tid() == 0) ++X.n_evals_full;
}

// _evaluate(u, l, hessian=False): x, g, q, G'l only (sensitivities optional)
template <bool SM>
DG_DEVN void eval_grad(Cta& c, SolveCtx& X, const double* u, const double* l_in, bool with_sens) {
  const Dims D = X.D; const EvalBuf E = X.W.E; DG_SH_EVAL(E);
  c.sync();
  c.lap(PH_OTHER);
  { double* lb = E.lbuf; DG_FOR(r, D.m) lb[r] = l_in[r]; }
  const double* l = E.lbuf;
  game_rollout<SM>(c, *X.G, D, u, X.x0, E.x, E.tmpS);
  c.sync();
  game_linearize<SM>(c, *X.G, D, u, E, false);
  c.sync();
  c.lap(PH_LIN_GRAD);
  game_constraints<SM>(c, *X.G, D, u, X.W.S.up, E.x, E.g, E.rowtab);
  game_costates<SM>(c, *X.G, D, E, l);
  if (with_sens) game_sens<SM>(c, D, E);
  c.sync();
  game_gradients<SM>(c, *X.G, D, E, u, X.W.S.up, l, X.P->merit_obj ? X.W.S.qs : nullptr);
  c.sync();
  c.lap(PH_ADJ_GRAD);
  if (c.tid() == 0) ++X.n_evals_grad;
}

// phi at the currently evaluated point:  1/2 |q+G'l|^2 + 1/2 (l.g)^2 + mu * sum(g - (s + alpha*ds))
template <bool SM>
DG_DEVN double merit_here(Cta& c, SolveCtx& X, const double* l, const double* s, const double* ds, double alpha,
                         double mu) {
  const Dims D = X.D; const EvalBuf E = X.W.E; DG_SH_EVAL(E);
  double p1 = 0.0, p2 = 0.0, p3 = 0.0;
  DG_FOR(i, D.n) { double d = E.q[i] + E.gtl[i]; p1 += d * d; }
  DG_FOR(r, D.m) { p2 += l[r] * E.g[r]; p3 += E.g[r] - (s[r] + (ds ? alpha * ds[r] : 0.0)); }
  c.sum3(p1, p2, p3);
  double dd = p1, lg = p2, vio = p3;
  double val = 0.5 * (dd + lg * lg);
  if (X.P->merit_l1) val += mu * vio;
  c.lap(PH_MERIT);
  return val;
}

// After a QP at the currently (fully) evaluated point (u_b, l_b): fills dl, s, ds, Gdu and returns
// phi, dphi (and mu when compute_mu) -- f_phi / f_dphi / _get_mu.
template <bool SM>
DG_DEVN void step_merit(Cta& c, SolveCtx& X, const double* l_b, const double* du, const double* l_hat,
                        double* dl, double* s, double* ds, bool compute_mu, double& mu, double& phi, double& dphi) {
  const Dims D = X.D; const EvalBuf E = X.W.E; DG_SH_EVAL(E); const SqpBuf S = X.W.S;
  const int n = D.n, m = D.m;
  c.lap(PH_OTHER);
  game_G_times<SM>(c, D, E, du, S.Gdu);
  DG_FOR(r, m) {
    dl[r] = l_hat[r] - l_b[r];
    double sv = E.g[r] < 0.0 ? E.g[r] : 0.0;
    s[r] = sv;
    ds[r] = E.g[r] + S.Gdu[r] - sv;
  }
  c.sync();
  game_GT_times<SM>(c, D, E, dl, S.tn2);
  // tn = Q du (raw, unsymmetrised Q -- DGSQP.py:416 passes Q_i)
  // (Q streams from global memory: warp per row, lanes along the row)
  for (int i = c.warp(); i < n; i += c.nwarps()) {
    const double* DG_RESTRICT Qi = E.Q + (size_t)i * n;
    double acc = 0.0;
    for (int j = c.lane(); j < n; j += c.wsz) acc += Qi[j] * du[j];
    acc = c.warp_sum(acc);
    if (c.lane() == 0) S.tn[i] = acc;
  }
  c.sync();
  double p1 = 0.0, p2 = 0.0, p3 = 0.0, p4 = 0.0, p5 = 0.0, p6 = 0.0;
  DG_FOR(i, n) { double d = E.q[i] + E.gtl[i]; p1 += d * d; p2 += d * (S.tn[i] + S.tn2[i]); }
  DG_FOR(r, m) {
    p3 += l_b[r] * E.g[r]; p4 += l_b[r] * S.Gdu[r]; p5 += dl[r] * E.g[r]; p6 += E.g[r] - s[r];
  }
  c.sum3(p1, p2, p3);
  c.sum3(p4, p5, p6);
  double dd = p1, dQ = p2, lg = p3, lGdu = p4, dlg = p5, vio = p6;
  double dstat = dQ + lg * (lGdu + dlg);
  if (compute_mu) {
    mu = 0.0;
    if (X.P->merit_l1 && vio > X.P->mu_vio_thresh) mu = fabs(dstat) / ((1.0 - 0.5) * vio);
  }
  phi = 0.5 * (dd + lg * lg);
  dphi = dstat;
  if (X.P->merit_l1) { phi += mu * vio; dphi -= mu * vio; }
  c.lap(PH_MERIT);
}

// _solve_qp at the currently evaluated point.  Result in W.Q.xq / W.Q.lam.  Returns 0 on success.
template <bool SM>
DG_DEVN int solve_qp_here(Cta& c, SolveCtx& X) {
  const Dims D = X.D;
  int nneg = nearest_pd<SM>(c, D.n, X.W.E.Q, X.W.B, X.P->eig_floor, X.P->reg, X.P->conv_approx != 0);
  if (c.tid() == 0) {
    if (nneg > X.n_neg_max) X.n_neg_max = nneg;
    if (nneg > 0) { ++X.n_qp_indef; X.n_neg_sum += nneg; }
  }
  int it = 0, na = 0;
  int st = qp_solve_gi<SM>(c, D, X.W.E, X.W.E.q, X.W.Q, X.W.B, &it, &na, X.P->qp_warm);
  if (c.tid() == 0) { X.n_gi_iters += it; X.n_act_sum += na; }
  return st;
}

// _line_search_3: base (u,du,l,dl,s,ds) with phi0/dphi0; result left in (u_c, l_c); returns phi_trial
template <bool SM>
DG_DEVN double line_search_3(Cta& c, SolveCtx& X, const double* u, const double* du, const double* l, const double* dl,
                             const double* s, const double* ds, double phi0, double dphi0, double mu) {
  const Dims D = X.D; const SqpBuf S = X.W.S;
  double alpha = 1.0, phi_t = 0.0;
  for (int i = 0; i < X.P->line_search_iters; ++i) {
    c.sync();
    DG_FOR(j, D.n) S.u_c[j] = u[j] + alpha * du[j];
    DG_FOR(r, D.m) S.l_c[r] = l[r] + alpha * dl[r];
    eval_grad<SM>(c, X, S.u_c, S.l_c, false);
    if (c.tid() == 0) ++X.n_ls_trials;
    phi_t = merit_here<SM>(c, X, S.l_c, s, ds, alpha, mu);
    if (phi_t <= phi0 + X.P->beta * alpha * dphi0) break;
    alpha *= X.P->tau;
  }
  return phi_t;
}

template <bool SM>
DG_DEV void vcopy(Cta& c, int len, double* dst, const double* src) { DG_FOR(i, len) dst[i] = src[i]; }

// _watchdog_line_search_4.  On return the accepted iterate is in (S.u, S.l); returns extra QP count.
template <bool SM>
DG_DEVN int watchdog_4(Cta& c, SolveCtx& X, double phi_k, double dphi_k, double mu) {
  const Dims D = X.D; const SqpBuf S = X.W.S; const SolverParams P = *X.P;
  const int n = D.n, m = D.m;
  int qp = 0;
  const double target = phi_k + P.beta * dphi_k;
  // relaxed (full) step
  c.sync();
  DG_FOR(j, n) S.u_c[j] = S.u[j] + S.du[j];
  DG_FOR(r, m) S.l_c[r] = S.l[r] + S.dl[r];
  eval_grad<SM>(c, X, S.u_c, S.l_c, false);
  double phi1 = merit_here<SM>(c, X, S.l_c, S.s, S.ds, 1.0, mu);
#ifdef DG_TRACE
  if (c.tid() == 0) printf("      full step phi1 %.12e  (accept %d)\n", phi1, (int)(phi1 <= target));
#endif
  if (phi1 <= target) { c.sync(); vcopy<SM>(c, n, S.u, S.u_c); vcopy<SM>(c, m, S.l, S.l_c); c.sync(); return qp; }
  bool fail = false;
  c.sync();
  vcopy<SM>(c, n, S.u_t, S.u_c); vcopy<SM>(c, m, S.l_t, S.l_c);
  for (int t = 0; t < P.t_hat; ++t) {
    eval_full<SM>(c, X, S.u_t, S.l_t);
    int st = solve_qp_here<SM>(c, X);
    ++qp;
    if (st != 0) { fail = true; break; }
    c.sync();
    vcopy<SM>(c, n, S.du_t, X.W.Q.xq);
    double mu_d = mu, ph, dph;
    step_merit<SM>(c, X, S.l_t, S.du_t, X.W.Q.lam, S.dl_t, S.s_t, S.ds_t, false, mu_d, ph, dph);
    c.sync();
    DG_FOR(j, n) S.u_c[j] = S.u_t[j] + S.du_t[j];
    DG_FOR(r, m) S.l_c[r] = X.W.Q.lam[r];
    eval_grad<SM>(c, X, S.u_c, S.l_c, false);
    double phi_n = merit_here<SM>(c, X, S.l_c, S.s_t, S.ds_t, 1.0, mu);
    if (phi_n > P.merit_max) break;
    if (phi_n <= target) { c.sync(); vcopy<SM>(c, n, S.u, S.u_c); vcopy<SM>(c, m, S.l, S.l_c); c.sync(); return qp; }
    c.sync();
    vcopy<SM>(c, n, S.u_t, S.u_c); vcopy<SM>(c, m, S.l_t, S.l_c);
  }
  // insist on merit decrease
  double phi_n = 0.0;
  {
    eval_full<SM>(c, X, S.u_t, S.l_t);
    int st = solve_qp_here<SM>(c, X);
    ++qp;
    if (st != 0) fail = true;
    else {
      c.sync();
      vcopy<SM>(c, n, S.du_t, X.W.Q.xq);
      double mu_d = mu, ph, dph;
      step_merit<SM>(c, X, S.l_t, S.du_t, X.W.Q.lam, S.dl_t, S.s_t, S.ds_t, false, mu_d, ph, dph);
      phi_n = line_search_3<SM>(c, X, S.u_t, S.du_t, S.l_t, S.dl_t, S.s_t, S.ds_t, ph, dph, mu);
    }
  }
  if (!fail) {
    if (phi_n <= target) { c.sync(); vcopy<SM>(c, n, S.u, S.u_c); vcopy<SM>(c, m, S.l, S.l_c); c.sync(); return qp; }
    else if (phi_n > phi_k) fail = true;
    else {
      c.sync();
      vcopy<SM>(c, n, S.u_t, S.u_c); vcopy<SM>(c, m, S.l_t, S.l_c);
      eval_full<SM>(c, X, S.u_t, S.l_t);
      int st = solve_qp_here<SM>(c, X);
      if (st != 0) {
        line_search_3<SM>(c, X, S.u, S.du, S.l, S.dl, S.s, S.ds, phi_k, dphi_k, mu);
        c.sync(); vcopy<SM>(c, n, S.u, S.u_c); vcopy<SM>(c, m, S.l, S.l_c); c.sync();
        return qp;
      }
      ++qp;
      c.sync();
      vcopy<SM>(c, n, S.du_t, X.W.Q.xq);
      double mu_d = mu, ph, dph;
      step_merit<SM>(c, X, S.l_t, S.du_t, X.W.Q.lam, S.dl_t, S.s_t, S.ds_t, false, mu_d, ph, dph);
      line_search_3<SM>(c, X, S.u_t, S.du_t, S.l_t, S.dl_t, S.s_t, S.ds_t, ph, dph, mu);
      c.sync(); vcopy<SM>(c, n, S.u, S.u_c); vcopy<SM>(c, m, S.l, S.l_c); c.sync();
      return qp;
    }
  }
  // fail: search along the original step
  line_search_3<SM>(c, X, S.u, S.du, S.l, S.dl, S.s, S.ds, phi_k, dphi_k, mu);
  c.sync(); vcopy<SM>(c, n, S.u, S.u_c); vcopy<SM>(c, m, S.l, S.l_c); c.sync();
  return qp;
}

struct SolveOut {
  double* u;      // n
  double* l;      // m
  double* x;      // (N+1)*nq
  double* cost;   // M
  double* cond;   // 3: p_feas, comp, stat
  int* num_iters; int* status; int* qp_solves;
  int* diag;      // DG_NDIAG work counters (may be null), see dgsqp_last_diag
  double* l_init; // m  (may be null): dual initialisation
  double* iter_log; int iter_cap;   // [iter_cap][DG_ITER_REC] per-iteration record (may be null), see dgsqp_last_iter_data
};

#define DG_ITER_REC 5     // p_feas, comp, stat, qp_solves, seconds
// iter_data.append(dict(cond, ..., qp_solves, it_time)) (DGSQP.py:445-452): one record per completed iteration
DG_DEV void iter_log_put(const Cta& c, const SolveOut& O, int it, double pf, double comp, double stat, int qp, double t_ns) {
  if (O.iter_log && c.tid() == 0 && it < O.iter_cap) {
    double* r = O.iter_log + (size_t)it * DG_ITER_REC;
    r[0] = pf; r[1] = comp; r[2] = stat; r[3] = (double)qp; r[4] = t_ns * 1e-9;
  }
}

// l_ws: optional dual warm start (nullptr = the reference's LSQR initialisation, DGSQP.py:312-326)
template <bool SM>
DG_DEVN void sqp_solve_v1(Cta& c, SolveCtx& X, const double* u_ws, const double* l_ws, const SolveOut& O) {
  const Dims D = X.D; const SqpBuf S = X.W.S; const EvalBuf E = X.W.E; DG_SH_EVAL(E); const SolverParams P = *X.P;
  const int n = D.n, m = D.m;
  if (c.tid() == 0) { X.n_evals_full = X.n_evals_grad = X.n_gi_iters = X.n_neg_max = X.n_qp_indef = X.n_neg_sum = X.n_act_sum = X.n_ls_trials = 0; }
  if (c.tid() == 0) X.W.Q.act[n] = 0;          // no previous active set yet (qp_solve_gi's warm start)
  game_row_table<SM>(c, D, E.rowtab);
  DG_FOR(j, n) S.u[j] = u_ws[j];
  DG_FOR(r, m) S.l[r] = 0.0;
  DG_FOR(j, D.nu) S.up[j] = 0.0;
  c.sync();
  // dual initialisation (DGSQP.py:320-326)
  if (l_ws) {
    DG_FOR(r, m) S.l[r] = l_ws[r];
    c.sync();
  } else {
    eval_grad<SM>(c, X, S.u, S.l, true);
    lsqr_dual_init<SM>(c, D, E, X.W.L, E.q, S.l);
    c.lap(PH_LSQR);
  }
  if (P.dbg_l0_perturb != 0.0) {
    DG_FOR(r, m) S.l[r] *= 1.0 + P.dbg_l0_perturb * (2.0 * (double)((r * 2654435761u) % 1000u) / 1000.0 - 1.0);
    c.sync();
  }
  if (O.l_init) { DG_FOR(r, m) O.l_init[r] = S.l[r]; }
  int sqp_it = 0, rel_its = 0, total_qp = 0, status = ST_MAX_IT;
  double p_feas = 0.0, comp = 0.0, stat = 0.0;
  const bool timed = P.time_limit_ns > 0.0 || O.iter_log != nullptr;
  const double t_start = timed && c.tid() == 0 ? dg_now_ns() : 0.0;
  if (O.iter_log) { for (int t = c.tid(); t < O.iter_cap * DG_ITER_REC; t += c.nt()) O.iter_log[t] = 0.0; }
  while (true) {
    const double t_it = timed && c.tid() == 0 ? dg_now_ns() : 0.0;
    const int qp_before = total_qp;
    eval_full<SM>(c, X, S.u, S.l);
    double a1 = -1e300, a2 = 0.0, a3 = 0.0;
    // NaN must not look like convergence (fmax drops NaNs): map it to +inf -> 'diverged'
    DG_FOR(r, m) {
      double gv = E.g[r], cv = fabs(gv * S.l[r]);
      a1 = fmax(a1, gv != gv ? 1e300 : gv); a2 = fmax(a2, cv != cv ? 1e300 : cv);
    }
    DG_FOR(j, n) { double dv = fabs(E.q[j] + E.gtl[j]); a3 = fmax(a3, dv != dv ? 1e300 : dv); }
    c.max3(a1, a2, a3);
    p_feas = fmax(0.0, a1); comp = a2; stat = a3;
    c.sync();
    vcopy<SM>(c, n, S.u_im1, S.u); vcopy<SM>(c, m, S.l_im1, S.l);
    if (stat > P.diverge_tol) { status = ST_DIVERGED; break; }
    if (p_feas < P.p_tol && comp < P.d_tol && stat < P.d_tol) { status = ST_CONV_ABS; break; }
    int st = solve_qp_here<SM>(c, X);
    ++total_qp;
    if (st != 0) { status = ST_QP_FAIL; break; }
    c.sync();
    vcopy<SM>(c, n, S.du, X.W.Q.xq);
    double mu = 0.0, phi_k, dphi_k;
    step_merit<SM>(c, X, S.l, S.du, X.W.Q.lam, S.dl, S.s, S.ds, true, mu, phi_k, dphi_k);
#ifdef DG_TRACE
    if (c.tid() == 0) printf("it %2d pf %.6e comp %.6e stat %.6e | mu %.12e phi_k %.12e dphi_k %.12e target %.12e\n", sqp_it, p_feas, comp, stat, mu, phi_k, dphi_k, phi_k + P.beta * dphi_k);
#endif
    if (P.nonmono_ls) total_qp += watchdog_4<SM>(c, X, phi_k, dphi_k, mu);
    else {
      line_search_3<SM>(c, X, S.u, S.du, S.l, S.dl, S.s, S.ds, phi_k, dphi_k, mu);
      c.sync(); vcopy<SM>(c, n, S.u, S.u_c); vcopy<SM>(c, m, S.l, S.l_c); c.sync();
    }
    double q1 = 0.0, q2 = 0.0;
    DG_FOR(j, n) { double d = S.u[j] - S.u_im1[j]; q1 += d * d; }
    DG_FOR(r, m) { double d = S.l[r] - S.l_im1[r]; q2 += d * d; }
    c.sum2(q1, q2);
    double nu_ = sqrt(q1), nl_ = sqrt(q2);
    iter_log_put(c, O, sqp_it, p_feas, comp, stat, total_qp - qp_before, timed && c.tid() == 0 ? dg_now_ns() - t_it : 0.0);
    if (nu_ < P.p_tol / 2 && nl_ < P.d_tol / 2) {
      ++rel_its;
      if (rel_its >= P.rel_tol_req && p_feas < P.p_tol) { status = ST_CONV_REL; break; }
    } else rel_its = 0;
    ++sqp_it;
    if (sqp_it >= P.sqp_iters) { status = ST_MAX_IT; break; }
    // time_limit (DGSQP.py:470-474), per instance on the device clock; the decision is taken by thread 0 and broadcast
    if (P.time_limit_ns > 0.0 && c.bcast0(c.tid() == 0 && dg_now_ns() - t_start > P.time_limit_ns)) { status = ST_TIME_LIMIT; break; }
  }
  // outputs: x_bar = evaluate_dynamics(u), costs f_J  (DGSQP.py:476-498)
  c.sync();
  game_rollout<SM>(c, *X.G, D, S.u, X.x0, E.x, E.tmpS);
  c.sync();
  DG_FOR(j, n) O.u[j] = S.u[j];
  DG_FOR(r, m) O.l[r] = S.l[r];
  DG_FOR(j, (D.N + 1) * D.nq) O.x[j] = E.x[j];
  game_costs<SM>(c, *X.G, D, S.u, S.up, E.x, O.cost);
  if (c.tid() == 0) {
    O.cond[0] = p_feas; O.cond[1] = comp; O.cond[2] = stat;
    *O.num_iters = sqp_it; *O.status = status; *O.qp_solves = total_qp;
    if (O.diag) {
      O.diag[0] = X.n_evals_full; O.diag[1] = X.n_evals_grad; O.diag[2] = X.n_gi_iters; O.diag[3] = X.n_neg_max;
      O.diag[4] = X.n_qp_indef; O.diag[5] = X.n_neg_sum; O.diag[6] = X.n_act_sum; O.diag[7] = X.n_ls_trials;
    }
  }
  c.sync();
}
