// QP sub-problem   min 1/2 du'H du + q'du   s.t.  G du <= -g    (DGSQP._solve_qp, DGSQP.py:232-266)
//
// The reference sends this to OSQP (polish=True) through CasADi's conic interface; when polish
// succeeds the result is the exact KKT point of the strictly convex QP.  Here that point is computed
// directly by the Goldfarb-Idnani dual active-set method, one QP per CTA: H = L L' (Cholesky),
// J = L^-T, unconstrained minimiser, then add the most violated constraint / drop blocking ones.
// G is applied matrix-free through the game's sensitivity rows (racing_game.cuh).
// Conventions: slack s_i = -g_i - G_i x >= 0, normal n_i = -G_i', multipliers lam >= 0,
// H x + q + G' lam = 0.
#pragma once
#include "racing_game.cuh"
#include "linalg.cuh"

#define DG_QP_FEAS_TOL 1e-10
#define DG_QP_DEP_TOL 1e-20

struct QpBuf {
  double* Jm;     // n*n
  double* Rm;     // n*n  (upper triangular factor of the active normals in J-coordinates)
  double* xq;     // n   primal iterate (du)
  double* dv;     // n   J' n_p
  double* zv;     // n   primal step direction
  double* rv;     // n   dual step direction (active part)
  double* npv;    // n   normal of the entering constraint
  double* hv;     // n   householder vector
  double* wv;     // n   scratch
  double* lam_act;// n
  double* sl;     // m   slacks
  double* lam;    // m   output multipliers
  int* act;       // n   active constraint ids
  int* is_act;    // m   flags
};

// returns 0 ok, 1 not PD, 2 infeasible, 3 iteration limit.  Output: Q.xq (du), Q.lam (l_hat).
// n_iter / n_active are optional diagnostics (uniform).
DG_DEVN int qp_solve_gi(Cta& c, const Dims& D, const EvalBuf& E, double* Hm, const double* qv, const QpBuf& Q,
                        int* n_iter_out, int* n_active_out) {
  const int n = D.n, m = D.m;
  if (!cholesky_lower(c, n, Hm)) return 1;
  tri_inverse_T(c, n, Hm, Q.Jm);
  // x = -J J' q
  DG_FOR(j, n) {
    double acc = 0.0;
    for (int i = 0; i <= j; ++i) acc += Q.Jm[i * n + j] * qv[i];
    Q.dv[j] = acc;
  }
  c.sync();
  DG_FOR(i, n) {
    double acc = 0.0;
    for (int j = i; j < n; ++j) acc += Q.Jm[i * n + j] * Q.dv[j];
    Q.xq[i] = -acc;
  }
  DG_FOR(r, m) { Q.is_act[r] = 0; Q.lam[r] = 0.0; }
  c.sync();
  int iq = 0, it = 0;
  const int max_iter = 10 * (n + m);
  int status = 0;
  while (true) {
    // slacks of all constraints
    game_G_times(c, D, E, Q.xq, Q.sl);
    double best = 0.0; int bi = 0x7fffffff;
    {
      double bv = 1e300; int bidx = 0x7fffffff;
      DG_FOR(r, m) {
        double sv = Q.is_act[r] ? 0.0 : -E.g[r] - Q.sl[r];
        if (sv < bv) { bv = sv; bidx = r; }
      }
      c.argmin(bv, bidx, best, bi);
    }
    if (!(best < -DG_QP_FEAS_TOL)) break;
    const int p = bi;
    game_G_row(c, D, E, p, Q.npv);                 // npv = G[p,:]  (normal is -npv)
    double lam_p = 0.0;
    bool added = false;
    while (!added) {
      if (++it > max_iter) { status = 3; break; }
      // d = J' n_p  (n_p = -G_p')
      double dd_all = 0.0, dd_tail = 0.0;
      DG_FOR(j, n) {
        double acc = 0.0;
        for (int i = 0; i < n; ++i) acc += Q.Jm[i * n + j] * Q.npv[i];
        acc = -acc;
        Q.dv[j] = acc;
        dd_all += acc * acc;
        if (j >= iq) dd_tail += acc * acc;
      }
      double zn = c.sum(dd_tail);
      double dall = c.sum(dd_all);
      // z = J[:, iq:] d[iq:]     r = R^-1 d[:iq]
      DG_FOR(i, n) {
        double acc = 0.0;
        for (int j = iq; j < n; ++j) acc += Q.Jm[i * n + j] * Q.dv[j];
        Q.zv[i] = acc;
      }
      if (c.tid == 0) {
        for (int i = iq - 1; i >= 0; --i) {
          double acc = Q.dv[i];
          for (int j = i + 1; j < iq; ++j) acc -= Q.Rm[i * n + j] * Q.rv[j];
          Q.rv[i] = acc / Q.Rm[i * n + i];
        }
      }
      c.sync();
      // dual step bound t1
      double t1 = 1e300; int ldrop = -1;
      for (int k = 0; k < iq; ++k) {
        double rk = Q.rv[k];
        if (rk > 0.0) { double tk = Q.lam_act[k] / rk; if (tk < t1) { t1 = tk; ldrop = k; } }
      }
      // primal step length t2
      double t2 = 1e300;
      if (zn > DG_QP_DEP_TOL * dall && zn > 0.0) {
        double part = 0.0;
        DG_FOR(i, n) part += Q.npv[i] * Q.xq[i];
        double gx = c.sum(part);
        double sp = -E.g[p] - gx;
        t2 = -sp / zn;
      }
      double t = t1 < t2 ? t1 : t2;
      if (t >= 1e300) { status = 2; break; }
      c.sync();                                   // all threads have read rv/lam_act before they change
      if (t2 < 1e300) { DG_FOR(i, n) Q.xq[i] += t * Q.zv[i]; }
      DG_FOR(k, iq) Q.lam_act[k] -= t * Q.rv[k];
      lam_p += t;
      c.sync();
      if (t2 <= t1) {
        // full step: add p.  Householder on d[iq:] -> (alpha, 0, ..., 0); J[:, iq:] <- J[:, iq:] P
        const int len = n - iq;
        double d0 = Q.dv[iq];
        double alpha = sqrt(zn);
        if (d0 > 0.0) alpha = -alpha;
        double vpart = 0.0;
        DG_FOR(j, len) {
          double vj = Q.dv[iq + j] - (j == 0 ? alpha : 0.0);
          Q.hv[j] = vj;
          vpart += vj * vj;
        }
        double vv = c.sum(vpart);
        if (vv > 0.0) {
          DG_FOR(i, n) {
            double acc = 0.0;
            for (int j = 0; j < len; ++j) acc += Q.Jm[i * n + iq + j] * Q.hv[j];
            Q.wv[i] = acc * (2.0 / vv);
          }
          c.sync();
          DG_FOR(j, len) {
            double vj = Q.hv[j];
            for (int i = 0; i < n; ++i) Q.Jm[i * n + iq + j] -= Q.wv[i] * vj;
          }
        }
        DG_FOR(i, iq) Q.Rm[i * n + iq] = Q.dv[i];
        if (c.tid == 0) {
          Q.Rm[iq * n + iq] = alpha;
          Q.act[iq] = p; Q.is_act[p] = 1; Q.lam_act[iq] = lam_p;
        }
        ++iq;
        added = true;
        c.sync();
      } else {
        // partial step: drop active constraint ldrop (Givens re-triangularisation), keep p
        if (c.tid == 0) {
          Q.is_act[Q.act[ldrop]] = 0;
          for (int k = ldrop; k < iq - 1; ++k) { Q.act[k] = Q.act[k + 1]; Q.lam_act[k] = Q.lam_act[k + 1]; }
        }
        // shift columns of R left: thread per row
        DG_FOR(i, iq) {
          for (int j = ldrop; j < iq - 1; ++j) Q.Rm[i * n + j] = Q.Rm[i * n + j + 1];
          Q.Rm[i * n + iq - 1] = 0.0;
        }
        c.sync();
        for (int j = ldrop; j < iq - 1; ++j) {
          double a = Q.Rm[j * n + j], b = Q.Rm[(j + 1) * n + j];
          double h = hypot(a, b);
          c.sync();                               // rotation parameters read before rows change
          if (h != 0.0) {
            double cs = a / h, sn = b / h;
            for (int col = j + c.tid; col < iq - 1; col += c.nt) {
              double r0 = Q.Rm[j * n + col], r1 = Q.Rm[(j + 1) * n + col];
              Q.Rm[j * n + col] = cs * r0 + sn * r1;
              Q.Rm[(j + 1) * n + col] = -sn * r0 + cs * r1;
            }
            DG_FOR(i, n) {
              double j0 = Q.Jm[i * n + j], j1 = Q.Jm[i * n + j + 1];
              Q.Jm[i * n + j] = cs * j0 + sn * j1;
              Q.Jm[i * n + j + 1] = -sn * j0 + cs * j1;
            }
          }
          c.sync();
        }
        --iq;
      }
    }
    if (status != 0) break;
  }
  if (status == 0) {
    c.sync();
    DG_FOR(k, iq) Q.lam[Q.act[k]] = Q.lam_act[k];
    c.sync();
  }
  if (n_iter_out) *n_iter_out = it;
  if (n_active_out) *n_active_out = iq;
  return status;
}
