// DG-SQP v2 step policy for ONE game instance, executed by one CTA.
//
// Restates DGSQP.solve of DGSQP/solvers/DGSQP_v2.py:322-652 with _solve_qp (:253-285, eigenvalue floor 1e-9 :1273,
// decaying regularisation), _get_mu (:683-707), load_checkpoint (:709-727), line_search (:729-760) and the
// 'stat_l1' merit (:1143-1161):  stat = q + G'l,  phi = 1/2 |stat|^2 + mu * sum(s),  s = max(0, g),
// dphi = stat'(Q du + G' dl) - mu * sum(s).  The SQP approximation (_evaluate) and the QP are the v1 kernels.
//
// Iteration records.  The reference keeps every IterationData and reloads iter_data[checkpoint_index] (m-step
// rejected, :533-542) or iter_data[min(checkpoint_index, len-1)] (QP failure, :449-453).  Only two records can ever
// be addressed: the checkpoint iteration's and the last appended one, so the device keeps exactly those
// (rec_ckpt in the *_t buffers, rec_prev in the r_* buffers) plus the running one (c_u, c_l, du, dl, s, ds).
#pragma once
#include "sqp_v1.cuh"

#define DG_V2_MEM_MAX 16

// merit pieces at the currently evaluated point (inputs u): the smooth part -- 1/2 |q + G'l|^2 for 'stat_l1', the sum
// of the agents' costs for 'sum_obj_l1' (:1143-1151) -- and sum(max(0, g))
template <bool SM>
DG_DEVN void v2_point_terms(Cta& c, SolveCtx& X, const double* u, double& base, double& vio) {
  const Dims D = X.D; const EvalBuf E = X.W.E; DG_SH_EVAL(E);
  double p1 = 0.0, p2 = 0.0;
  if (X.P->merit_obj) {
    c.sync();
    game_costs<SM>(c, *X.G, D, u, X.W.S.up, E.x, E.cf);         // cf: scratch of the G' products, dead here
    c.sync();
    if (c.tid() == 0) for (int a = 0; a < D.M; ++a) p1 += E.cf[a];
  } else {
    DG_FOR(i, D.n) { double d = E.q[i] + E.gtl[i]; p1 += 0.5 * d * d; }
  }
  DG_FOR(r, D.m) { double gv = E.g[r]; p2 += gv > 0.0 ? gv : 0.0; }
  c.sum2(p1, p2);
  base = p1; vio = p2;
  c.lap(PH_MERIT);
}

// stat'(Q du + G' dl) at the currently (fully) evaluated point, with the raw game Hessian Q
template <bool SM>
DG_DEVN double v2_dstat(Cta& c, SolveCtx& X, const double* du, const double* dl) {
  const Dims D = X.D; const EvalBuf E = X.W.E; DG_SH_EVAL(E); const SqpBuf S = X.W.S;
  const int n = D.n;
  if (X.P->merit_obj) {
    // d(sum of costs) along du (:1150-1151); the multipliers do not enter
    double p = 0.0;
    DG_FOR(i, n) p += S.qs[i] * du[i];
    p = c.sum(p);
    c.lap(PH_MERIT);
    return p;
  }
  game_GT_times<SM>(c, D, E, dl, S.tn2);
  for (int i = c.warp(); i < n; i += c.nwarps()) {
    const double* DG_RESTRICT Qi = E.Q + (size_t)i * n;
    double acc = 0.0;
    for (int j = c.lane(); j < n; j += c.wsz) acc += Qi[j] * du[j];
    acc = c.warp_sum(acc);
    if (c.lane() == 0) S.tn[i] = acc;
  }
  c.sync();
  double p = 0.0;
  DG_FOR(i, n) p += (E.q[i] + E.gtl[i]) * (S.tn[i] + S.tn2[i]);
  p = c.sum(p);
  c.lap(PH_MERIT);
  return p;
}

// after a successful QP at the evaluated point (u, l_b):  dl = l_hat - l_b,  s = max(0, g),
// ds = max(0, g + G du) - s,  |(du, dl)|,  sum(s)
template <bool SM>
DG_DEVN void v2_step_vectors(Cta& c, SolveCtx& X, const double* l_b, const double* du, const double* l_hat,
                             double* dl, double* s, double* ds, double& step_norm, double& vio) {
  const Dims D = X.D; const EvalBuf E = X.W.E; DG_SH_EVAL(E); const SqpBuf S = X.W.S;
  c.lap(PH_OTHER);
  game_G_times<SM>(c, D, E, du, S.Gdu);
  double p1 = 0.0, p2 = 0.0, p3 = 0.0;
  DG_FOR(r, D.m) {
    const double dv = l_hat[r] - l_b[r];
    dl[r] = dv;
    const double gv = E.g[r], sv = gv > 0.0 ? gv : 0.0, gn = gv + S.Gdu[r];
    s[r] = sv;
    ds[r] = (gn > 0.0 ? gn : 0.0) - sv;
    p1 += dv * dv; p3 += sv;
  }
  DG_FOR(j, D.n) p2 += du[j] * du[j];
  c.sum3(p1, p2, p3);
  step_norm = sqrt(p1 + p2); vio = p3;
  c.lap(PH_MERIT);
}

// line_search (:729-760) from (S.u, S.l) along (S.du, S.dl) with slack S.s; the last trial is left in (S.u_c, S.l_c).
// Returns phi of the last trial with mu = 1 (what the reference appends to the merit memory).
template <bool SM>
DG_DEVN double v2_line_search(Cta& c, SolveCtx& X, double mu, double mem_max) {
  const Dims D = X.D; const SqpBuf S = X.W.S; const SolverParams P = *X.P;
  double phi0 = 0.0, dphi0 = 0.0;
  if (P.armijo) {
    eval_full<SM>(c, X, S.u, S.l);
    double base0, vio0;
    v2_point_terms<SM>(c, X, S.u, base0, vio0);
    double ssum = 0.0;
    DG_FOR(r, D.m) ssum += S.s[r];
    ssum = c.sum(ssum);
    phi0 = base0 + mu * ssum;
    dphi0 = v2_dstat<SM>(c, X, S.du, S.dl) - mu * vio0;
  }
  double a = 1.0, phi1 = 0.0;
  for (int it = 0; it < P.line_search_iters; ++it) {
    c.sync();
    DG_FOR(j, D.n) S.u_c[j] = S.u[j] + a * S.du[j];
    DG_FOR(r, D.m) S.l_c[r] = S.l[r] + a * S.dl[r];
    eval_grad<SM>(c, X, S.u_c, S.l_c, false);
    if (c.tid() == 0) ++X.n_ls_trials;
    double base, vio;
    v2_point_terms<SM>(c, X, S.u_c, base, vio);
    phi1 = base + vio;
    const double ref = P.armijo ? phi0 + P.sigma * a * dphi0 : (1.0 - P.sigma * a) * mem_max;
    if (base + mu * vio <= ref) break;
    a *= P.tau;
  }
  return phi1;
}

// one iteration record = (u, du, l, dl, s, ds) + mu
struct V2Rec { double *u, *du, *l, *dl, *s, *ds; };

template <bool SM>
DG_DEV void v2_copy_rec(Cta& c, const Dims& D, const V2Rec& dst, const V2Rec& src) {
  DG_FOR(j, D.n) { dst.u[j] = src.u[j]; dst.du[j] = src.du[j]; }
  DG_FOR(r, D.m) { dst.l[r] = src.l[r]; dst.dl[r] = src.dl[r]; dst.s[r] = src.s[r]; dst.ds[r] = src.ds[r]; }
}

template <bool SM>
DG_DEVN void sqp_solve_v2(Cta& c, SolveCtx& X, const double* u_ws, const double* l_ws, const SolveOut& O) {
  const Dims D = X.D; const SqpBuf S = X.W.S; const EvalBuf E = X.W.E; DG_SH_EVAL(E); const SolverParams P = *X.P;
  const int n = D.n, m = D.m;
  if (c.tid() == 0) { X.n_evals_full = X.n_evals_grad = X.n_gi_iters = X.n_neg_max = X.n_qp_indef = X.n_neg_sum = X.n_act_sum = X.n_ls_trials = 0; }
  if (c.tid() == 0) X.W.Q.act[n] = 0;          // no previous active set yet (qp_solve_gi's warm start)
  game_row_table<SM>(c, D, E.rowtab);
  DG_FOR(j, n) S.u[j] = u_ws[j];
  DG_FOR(r, m) S.l[r] = 0.0;
  DG_FOR(j, D.nu) S.up[j] = X.up_in ? X.up_in[j] : 0.0;          // u_prev of the solver object (:328)
  c.sync();
  // dual initialisation (:333-337) and the first entry of the merit memory (:342-343)
  eval_grad<SM>(c, X, S.u, S.l, true);
  if (l_ws) { DG_FOR(r, m) S.l[r] = l_ws[r]; c.sync(); }
  else { lsqr_dual_init<SM>(c, D, E, X.W.L, E.q, S.l); c.lap(PH_LSQR); }
  if (O.l_init) { DG_FOR(r, m) O.l_init[r] = S.l[r]; }
  double mem[DG_V2_MEM_MAX];
  int mem_len = 0, mem_pos = 0;
  const int mem_cap = P.nms_memory < 1 ? 1 : (P.nms_memory > DG_V2_MEM_MAX ? DG_V2_MEM_MAX : P.nms_memory);
  {
    // phi(u, l0) needs G'l0: gradient-only evaluation at (u, l0)
    eval_grad<SM>(c, X, S.u, S.l, false);
    double base, vio;
    v2_point_terms<SM>(c, X, S.u, base, vio);
    mem[0] = base + vio; mem_len = 1; mem_pos = 1 % mem_cap;
  }
  c.sync();
  vcopy<SM>(c, n, S.u_im1, S.u); vcopy<SM>(c, m, S.l_im1, S.l);
  const V2Rec cur = {S.c_u, S.du, S.c_l, S.dl, S.s, S.ds};
  const V2Rec ckp = {S.u_t, S.du_t, S.l_t, S.dl_t, S.s_t, S.ds_t};
  const V2Rec prv = {S.r_u, S.r_du, S.r_l, S.r_dl, S.r_s, S.r_ds};
  double mu = 0.0, mu_ckp = 0.0, mu_prv = 0.0;
  double reg = P.reg, delta = 0.0, ck_delta = 0.0, ck_reg = P.reg;
  int ck_counter = 0, ck_index = 0;
  int sqp_it = 0, m_step_it = 0, rel_its = 0, total_qp = 0, status = ST_MAX_IT;
  bool finished = false;
  double p_feas = 0.0, comp = 0.0, stat = 0.0;
  const bool timed = P.time_limit_ns > 0.0 || O.iter_log != nullptr;
  const double t_start = timed && c.tid() == 0 ? dg_now_ns() : 0.0;
  if (O.iter_log) { for (int t = c.tid(); t < O.iter_cap * DG_ITER_REC; t += c.nt()) O.iter_log[t] = 0.0; }
  while (true) {
    const double t_it = timed && c.tid() == 0 ? dg_now_ns() : 0.0;
    c.sync();
    vcopy<SM>(c, n, S.c_u, S.u); vcopy<SM>(c, m, S.c_l, S.l);
    eval_full<SM>(c, X, S.u, S.l);
    double a1 = -1e300, a2 = 0.0, a3 = 0.0;
    DG_FOR(r, m) {
      double gv = E.g[r], cv = fabs(gv * S.l[r]);
      a1 = fmax(a1, gv != gv ? 1e300 : gv); a2 = fmax(a2, cv != cv ? 1e300 : cv);
    }
    DG_FOR(j, n) { double dv = fabs(E.q[j] + E.gtl[j]); a3 = fmax(a3, dv != dv ? 1e300 : dv); }
    c.max3(a1, a2, a3);
    p_feas = fmax(0.0, a1); comp = a2; stat = a3;
    // the three tests run in the reference's order; a later one overrides the message of an earlier one (:394-411)
    if (stat > P.diverge_tol) { status = ST_DIVERGED; finished = true; }
    if (p_feas < P.p_tol && comp < P.d_tol && stat < P.d_tol) { status = ST_CONV_ABS; finished = true; }
    if (m_step_it >= P.sqp_iters) { status = ST_MAX_IT; finished = true; }
    // nms = False: no step is ever an m-step, so the reference's only bounds are convergence and time_limit (with the
    // default time_limit = None it never returns for an instance that does not settle).  A persistent kernel cannot
    // spin on one instance: without nms the total iteration count is capped at sqp_iters as well.
    if (!P.nms && sqp_it >= P.sqp_iters) { status = ST_MAX_IT; finished = true; }
    // time_limit_exceeded (DGSQP_v2.py:412), per instance on the device clock
    if (P.time_limit_ns > 0.0 && c.bcast0(c.tid() == 0 && dg_now_ns() - t_start > P.time_limit_ns)) { status = ST_TIME_LIMIT; finished = true; }
    if (finished) break;

    const bool is_ckpt_iter = sqp_it == ck_index;
    int nneg = nearest_pd<SM>(c, n, E.Q, X.W.B, P.eig_floor, reg, true);
    if (c.tid() == 0) { if (nneg > X.n_neg_max) X.n_neg_max = nneg; if (nneg > 0) { ++X.n_qp_indef; X.n_neg_sum += nneg; } }
    int gi_it = 0, gi_na = 0;
    const int qp_st = qp_solve_gi<SM>(c, D, E, E.q, X.W.Q, X.W.B, &gi_it, &gi_na, P.qp_warm);
    if (c.tid() == 0) { X.n_gi_iters += gi_it; X.n_act_sum += gi_na; }
    ++total_qp;
    bool d_step = false, m_step = false;
    int rec_src = 0;                        // record appended for this iteration: 0 running, 1 checkpoint's, 2 previous one's
    if (qp_st != 0) {
      if (!P.nms || sqp_it == 0) { status = ST_QP_FAIL; break; }
      m_step = true;
      // u, du, l, dl, s, ds, mu <- iter_data[min(checkpoint_index, len - 1)]
      c.sync();
      if (ck_index <= sqp_it - 1) { v2_copy_rec<SM>(c, D, cur, ckp); mu = mu_ckp; rec_src = 1; }
      else { v2_copy_rec<SM>(c, D, cur, prv); mu = mu_prv; rec_src = 2; }
      c.sync();
      vcopy<SM>(c, n, S.u, S.c_u); vcopy<SM>(c, m, S.l, S.c_l);
      c.sync();
    } else {
      c.sync();
      vcopy<SM>(c, n, S.du, X.W.Q.xq);
      double step_norm, vio;
      v2_step_vectors<SM>(c, X, S.l, S.du, X.W.Q.lam, S.dl, S.s, S.ds, step_norm, vio);
      if (sqp_it == 0) { delta = 20.0 * step_norm; ck_delta = delta; }            // nms_initial_step_size_factor (:212)
      if (P.nms) {
        if (ck_counter >= P.nms_frequency) m_step = true;
        else if (step_norm < delta) d_step = true;
        else m_step = true;
      }
      if (P.has_merit_parameter) mu = P.merit_parameter;
      else {
        const double dst = v2_dstat<SM>(c, X, S.du, S.dl);
        mu = vio > P.mu_vio_thresh ? fabs(dst) / ((1.0 - 0.5) * vio) : 0.0;
      }
    }
    if (d_step) {
      c.sync();
      DG_FOR(j, n) S.u[j] += S.du[j];
      DG_FOR(r, m) S.l[r] += S.dl[r];
      delta *= P.gamma;
      ++ck_counter;
    }
    double phi = 0.0;
    bool moved = false;                     // an m-step or a plain line-search step updates the reference iterate
    if (m_step) {
      ++m_step_it;
      c.sync();
      DG_FOR(j, n) S.u_c[j] = S.u[j] + S.du[j];
      DG_FOR(r, m) S.l_c[r] = S.l[r] + S.dl[r];
      eval_grad<SM>(c, X, S.u_c, S.l_c, false);
      double base, vio;
      v2_point_terms<SM>(c, X, S.u_c, base, vio);
      phi = base + vio;
      double mem_max = mem[0];
      for (int t = 1; t < mem_len; ++t) mem_max = fmax(mem_max, mem[t]);
      if (!(phi <= (1.0 - P.sigma) * mem_max)) {
        if (ck_index <= sqp_it - 1) {
          c.sync();
          v2_copy_rec<SM>(c, D, cur, ckp); mu = mu_ckp; rec_src = 1;
          c.sync();
          vcopy<SM>(c, n, S.u, S.c_u); vcopy<SM>(c, m, S.l, S.c_l);
          delta = ck_delta; reg = ck_reg;
        }
        phi = v2_line_search<SM>(c, X, mu, mem_max);
      }
      c.sync();
      vcopy<SM>(c, n, S.u, S.u_c); vcopy<SM>(c, m, S.l, S.l_c);
      moved = true;
    }
    if (!d_step && !m_step) {
      double mem_max = mem[0];
      for (int t = 1; t < mem_len; ++t) mem_max = fmax(mem_max, mem[t]);
      phi = v2_line_search<SM>(c, X, mu, mem_max);
      c.sync();
      vcopy<SM>(c, n, S.u, S.u_c); vcopy<SM>(c, m, S.l, S.l_c);
      moved = true;
    }
    if (moved) {
      c.sync();
      double q1 = 0.0, q2 = 0.0;
      DG_FOR(j, n) { double d = S.u[j] - S.u_im1[j]; q1 += d * d; }
      DG_FOR(r, m) { double d = S.l[r] - S.l_im1[r]; q2 += d * d; }
      c.sum2(q1, q2);
      if (sqrt(q1) < P.p_tol && sqrt(q2) < P.d_tol) {
        ++rel_its;
        if (rel_its >= P.rel_tol_req && p_feas < P.p_tol) { status = ST_CONV_REL; finished = true; }   // acted upon next turn (:415)
      } else rel_its = 0;
      vcopy<SM>(c, n, S.u_im1, S.u); vcopy<SM>(c, m, S.l_im1, S.l);
      reg *= P.reg_decay;
      mem[mem_pos] = phi; mem_pos = (mem_pos + 1) % mem_cap; if (mem_len < mem_cap) ++mem_len;
      if (m_step) { ck_counter = 0; ck_delta = delta; ck_reg = reg; ck_index = sqp_it + 1; }
    }
    // iter_data.append(_data)
    c.sync();
    if (rec_src == 0) { v2_copy_rec<SM>(c, D, prv, cur); mu_prv = mu; }
    else if (rec_src == 1) { v2_copy_rec<SM>(c, D, prv, ckp); mu_prv = mu_ckp; }
    if (is_ckpt_iter) { c.sync(); v2_copy_rec<SM>(c, D, ckp, prv); mu_ckp = mu_prv; }
    // IterationData: primal_feasibility, complementarity, stationarity, qp_solutions, iteration_time (DGSQP_v2.py:31-52)
    iter_log_put(c, O, sqp_it, p_feas, comp, stat, 1, timed && c.tid() == 0 ? dg_now_ns() - t_it : 0.0);
    ++sqp_it;
  }
  // outputs (:604-647)
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
      O.diag[4] = X.n_qp_indef; O.diag[5] = X.n_neg_sum; O.diag[6] = m_step_it; O.diag[7] = X.n_ls_trials;
    }
  }
  c.sync();
}
