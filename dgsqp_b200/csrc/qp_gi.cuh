// QP sub-problem   min 1/2 du'H du + q'du   s.t.  G du <= -g    (DGSQP._solve_qp, DGSQP.py:232-266)
//
// The reference sends this to OSQP (polish=True) through CasADi's conic interface; when polish
// succeeds the result is the exact KKT point of the strictly convex QP.  Here that point is computed
// directly by the Goldfarb-Idnani dual active-set method, one QP per CTA: H = L L' (Cholesky),
// J = L^-T, unconstrained minimiser, then add the most violated constraint / drop blocking ones.
// G is applied matrix-free through the game's sensitivity rows (racing_game.cuh).
// Conventions: slack s_i = -g_i - G_i x >= 0, normal n_i = -G_i', multipliers lam >= 0,
// H x + q + G' lam = 0.
//
// Storage: Y = J' (row-major), i.e. J(i,j) = Y[j*n+i].  Y starts as L^-1 (tri_inverse) and every
// per-iteration O(n^2) operation touches it with thread = column of Y (coalesced): z = J2 d2, the
// Householder update of J's trailing columns and the Givens rotations of a drop.  Only d = J'n needs
// rows of Y; it is computed warp-per-row.
#pragma once
#include "racing_game.cuh"
#include "linalg.cuh"

#define DG_QP_FEAS_TOL 1e-10
#define DG_QP_DEP_TOL 1e-20

struct QpBuf {
  double* Y;      // n*n   J transposed
  double* Rm;     // n*n   upper triangular factor of the active normals in J-coordinates
  double* xq;     // n   primal iterate (du)
  double* dv;     // n   J' n_p
  double* zv;     // n   primal step direction
  double* rv;     // n   dual step direction (active part)
  double* npv;    // n   G row of the entering constraint (normal is -npv)
  double* lam_act;// n
  double* sl;     // m   G x
  double* lam;    // m   output multipliers
  int* act;       // n   active constraint ids
  int* is_act;    // m   flags
};

// returns 0 ok, 1 not PD, 2 infeasible, 3 iteration limit.  Output: Q.xq (du), Q.lam (l_hat).
DG_DEVN int qp_solve_gi(Cta& c, const Dims& D, const EvalBuf& E, double* DG_RESTRICT Hm, const double* DG_RESTRICT qv,
                        const QpBuf& Q, const LinBuf& B, int* n_iter_out, int* n_active_out) {
  const int n = D.n, m = D.m;
  double* DG_RESTRICT Y = Q.Y;
  if (!cholesky_lower(c, n, Hm, B.sp)) return 1;
  c.lap(PH_CHOL);
  tri_inverse(c, n, Hm, Y);
  // x = -J J' q = -Y' (Y q):   t = Y q (warp per row), x_i = -sum_j Y[j][i] t_j (thread per column)
  for (int j = c.warp; j < n; j += c.nwarps) {
    const double* DG_RESTRICT Yj = Y + j * n;
    double p = 0.0;
    for (int i = c.lane; i <= j; i += c.wsz) p += Yj[i] * qv[i];
    p = c.warp_sum(p);
    if (c.lane == 0) Q.dv[j] = p;
  }
  DG_FOR(r, m) { Q.is_act[r] = 0; Q.lam[r] = 0.0; }
  c.sync();
  DG_FOR(i, n) {
    double a0 = 0.0, a1 = 0.0;
    int j = i;
    for (; j + 2 <= n; j += 2) { a0 += Y[j * n + i] * Q.dv[j]; a1 += Y[(j + 1) * n + i] * Q.dv[j + 1]; }
    for (; j < n; ++j) a0 += Y[j * n + i] * Q.dv[j];
    Q.xq[i] = -(a0 + a1);
  }
  c.lap(PH_TRINV);
  int iq = 0, it = 0;
  const int max_iter = 10 * (n + m);
  int status = 0;
  while (true) {
    // slacks of all constraints, most violated one
    game_G_times(c, D, E, Q.xq, Q.sl);
    double best; int bi;
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
      // d = J' n_p = -Y npv   (warp per row);  also npv . x
      double dd_tail = 0.0, dd_all = 0.0, gx = 0.0;
      for (int j = c.warp; j < n; j += c.nwarps) {
        const double* DG_RESTRICT Yj = Y + j * n;
        double pp = 0.0;
        for (int i = c.lane; i < n; i += c.wsz) pp += Yj[i] * Q.npv[i];
        pp = -c.warp_sum(pp);
        if (c.lane == 0) {
          Q.dv[j] = pp;
          dd_all += pp * pp;
          if (j >= iq) dd_tail += pp * pp;
        }
      }
      DG_FOR(i, n) gx += Q.npv[i] * Q.xq[i];
      c.sum3(dd_tail, dd_all, gx);
      const double zn = dd_tail, dall = dd_all;
      // z = J[:, iq:] d[iq:]  (thread per column of Y)     r = R^-1 d[:iq]  (thread 0)
      DG_FOR(i, n) {
        double a0 = 0.0, a1 = 0.0;
        int j = iq;
        for (; j + 2 <= n; j += 2) { a0 += Y[j * n + i] * Q.dv[j]; a1 += Y[(j + 1) * n + i] * Q.dv[j + 1]; }
        for (; j < n; ++j) a0 += Y[j * n + i] * Q.dv[j];
        Q.zv[i] = a0 + a1;
      }
      if (c.tid == c.nt - 1) {
        for (int i = iq - 1; i >= 0; --i) {
          double acc = Q.dv[i];
          for (int j = i + 1; j < iq; ++j) acc -= Q.Rm[i * n + j] * Q.rv[j];
          Q.rv[i] = acc / Q.Rm[i * n + i];
        }
      }
      c.sync();
      // dual step bound t1, primal step length t2
      double t1 = 1e300; int ldrop = -1;
      for (int k = 0; k < iq; ++k) {
        double rk = Q.rv[k];
        if (rk > 0.0) { double tk = Q.lam_act[k] / rk; if (tk < t1) { t1 = tk; ldrop = k; } }
      }
      double t2 = 1e300;
      if (zn > DG_QP_DEP_TOL * dall && zn > 0.0) t2 = (E.g[p] + gx) / zn;      // -s_p / |d2|^2
      const double t = t1 < t2 ? t1 : t2;
      if (t >= 1e300) { status = 2; break; }
      c.sync();                                   // rv / lam_act read by everyone before they change
      if (t2 < 1e300) { DG_FOR(i, n) Q.xq[i] += t * Q.zv[i]; }
      DG_FOR(k, iq) Q.lam_act[k] -= t * Q.rv[k];
      lam_p += t;
      if (t2 <= t1) {
        // full step: add p.  Householder P on d[iq:] -> (alpha, 0, ..); J[:, iq:] <- J[:, iq:] P, and
        // J2 v = z - alpha J[:, iq] needs no extra product.
        const int len = n - iq;
        const double d0 = Q.dv[iq];
        double alpha = sqrt(zn);
        if (d0 > 0.0) alpha = -alpha;
        const double vv = 2.0 * (zn - alpha * d0);
        if (vv > 0.0) {
          const double sc = 2.0 / vv;
          DG_FOR(i, n) {
            const double wi = (Q.zv[i] - alpha * Y[iq * n + i]) * sc;
            double* DG_RESTRICT col = Y + iq * n + i;
            col[0] -= (d0 - alpha) * wi;
            int j = 1;
            for (; j + 4 <= len; j += 4) {
              double y0 = col[(j + 0) * n], y1 = col[(j + 1) * n], y2 = col[(j + 2) * n], y3 = col[(j + 3) * n];
              y0 -= Q.dv[iq + j + 0] * wi; y1 -= Q.dv[iq + j + 1] * wi;
              y2 -= Q.dv[iq + j + 2] * wi; y3 -= Q.dv[iq + j + 3] * wi;
              col[(j + 0) * n] = y0; col[(j + 1) * n] = y1; col[(j + 2) * n] = y2; col[(j + 3) * n] = y3;
            }
            for (; j < len; ++j) col[j * n] -= Q.dv[iq + j] * wi;
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
        c.sync();
        // partial step: drop active constraint ldrop (Givens re-triangularisation), keep p
        if (c.tid == 0) {
          Q.is_act[Q.act[ldrop]] = 0;
          for (int k = ldrop; k < iq - 1; ++k) { Q.act[k] = Q.act[k + 1]; Q.lam_act[k] = Q.lam_act[k + 1]; }
        }
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
            DG_FOR(i, n) {                        // columns j, j+1 of J = rows j, j+1 of Y
              double j0 = Y[j * n + i], j1 = Y[(j + 1) * n + i];
              Y[j * n + i] = cs * j0 + sn * j1;
              Y[(j + 1) * n + i] = -sn * j0 + cs * j1;
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
  c.lap(PH_GI);
  return status;
}
