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
// Storage: Y = J' (row-major, leading dimension ld), i.e. J(i,j) = Y[j*ld+i].  Y starts as L^-1 (tri_inverse).
// Every O(n^2) operation of an iteration is spread over the whole CTA: products with J' run warp-per-row, products
// with J and the Householder update of J's trailing columns use the 2D decomposition of cta.cuh (consecutive
// threads = consecutive columns of Y, column groups interleave the rows) with partial sums combined through
// B.part; the triangular solve with R runs on one warp while the others combine the partial sums.
#pragma once
#include "game.cuh"
#include "linalg.cuh"

#define DG_QP_FEAS_TOL 1e-10
#define DG_QP_DEP_TOL 1e-20

struct QpBuf {
  // Y = J' lives in LinBuf::matB, R (upper triangular factor of the active normals in J-coordinates) in
  // LinBuf::matA once the Cholesky factor stored there has been inverted
  double* xq;     // n   primal iterate (du)
  double* dv;     // n   J' n_p
  double* zv;     // n   primal step direction
  double* rv;     // n   dual step direction (active part)
  double* npv;    // n   G row of the entering constraint (normal is -npv); reused for the Householder vector
  double* lam_act;// n
  double* sl;     // m   G x
  double* lam;    // m   output multipliers
  int* act;       // n   active constraint ids
  int* is_act;    // m   flags
};

#define DG_SH_QP(Q) do { DG_ASSUME_SHARED((Q).xq); DG_ASSUME_SHARED((Q).dv); DG_ASSUME_SHARED((Q).zv); DG_ASSUME_SHARED((Q).rv); \
  DG_ASSUME_SHARED((Q).npv); DG_ASSUME_SHARED((Q).lam_act); DG_ASSUME_SHARED((Q).sl); DG_ASSUME_SHARED((Q).act); DG_ASSUME_SHARED((Q).is_act); } while (0)

// out[i] = scale * sum_{j in [j0, n)} Y[j][i] * d[j],  i < n.  All threads; ends with a barrier.
template <bool SM>
DG_DEV void gi_cols_times(Cta& c, int n, int ld, const double* DG_RESTRICT Y, const double* DG_RESTRICT d, int j0,
                          double* DG_RESTRICT part, double* DG_RESTRICT out, double scale) {
  // (inlined: the caller carries the address-space hints)
  const Split2 sp = split2(c, n);
  for (int i = sp.i0; i < n; i += sp.istep) {
    double a0 = 0.0, a1 = 0.0;
    int j = j0 + sp.g;
    for (; j + sp.G < n; j += 2 * sp.G) { a0 += Y[j * ld + i] * d[j]; a1 += Y[(j + sp.G) * ld + i] * d[j + sp.G]; }
    if (j < n) a0 += Y[j * ld + i] * d[j];
    if (sp.G == 1) out[i] = scale * (a0 + a1);
    else part[sp.g * sp.istep + i] = a0 + a1;
  }
  c.sync();
  if (sp.G > 1) {
    if (sp.g == 0) {
      for (int i = sp.i0; i < n; i += sp.istep) {
        double acc = part[i];
        for (int g = 1; g < sp.G; ++g) acc += part[g * sp.istep + i];
        out[i] = scale * acc;
      }
    }
    c.sync();
  }
}

#ifdef DG_QP_WARM_START
// Warm start of the dual active-set method from the active set of the previous QP of this instance (SolverParams::qp_warm,
// off by default; successive SQP iterations share most of their active set).  Starting from the unconstrained minimiser
// x0 in Q.xq and Y = L^-1:
//   (1) the previous constraints are added to the factorisation one by one (same Householder update as a full step,
//       linearly dependent ones are skipped) without the step-length logic;
//   (2) the equality-constrained minimiser on that set follows from two triangular solves:
//       r = g_W + G_W x0,  t = R^-T r,  lam = R^-1 t,  x = x0 + J1 t;
//   (3) constraints whose multiplier comes out negative are dropped (most negative first, Givens re-triangularisation)
//       and (2) is repeated, until (x, lam >= 0, W) is an S-pair from which the method continues unchanged.
// The QP is strictly convex, so the solution is the same as from a cold start (up to rounding); only the path differs.
// Returns the number of active constraints; Q.act / Q.is_act / Q.lam_act / Q.xq are set accordingly.
// Prototype: plain CTA-wide loops, not tuned.
template <bool SM>
DG_DEVN int gi_warm_start(Cta& c, const Dims& D_, const EvalBuf& E_, const QpBuf& Q_, const LinBuf& B_, int nprev) {
  const EvalBuf E = E_; const QpBuf Q = Q_; const LinBuf B = B_; const Dims D = D_;
  const int n = D.n, ld = B.ld;
  double* Y = B.matB;
  double* Rm = B.matA;
  int iq = 0;
  for (int k = 0; k < nprev; ++k) {
    const int p = Q.act[k];
    c.sync();
    if (p < 0 || p >= D.m) continue;
    game_G_row<SM>(c, D, E, p, Q.npv);
    for (int j = c.warp(); j < n; j += c.nwarps()) {
      double acc = 0.0;
      for (int i = c.lane(); i < n; i += c.wsz) acc += Y[j * ld + i] * Q.npv[i];
      acc = c.warp_sum(acc);
      if (c.lane() == 0) Q.dv[j] = -acc;
    }
    c.sync();
    double zn = 0.0, dall = 0.0;
    DG_FOR(j, n) { const double dj = Q.dv[j]; dall += dj * dj; if (j >= iq) zn += dj * dj; }
    c.sum2(zn, dall);
    if (!(zn > DG_QP_DEP_TOL * dall && zn > 0.0)) continue;          // linearly dependent on the ones already in
    const int len = n - iq;
    const double d0 = Q.dv[iq];
    double alpha = sqrt(zn);
    if (d0 > 0.0) alpha = -alpha;
    const double vv = 2.0 * (zn - alpha * d0);
    c.sync();
    if (vv > 0.0) {
      const double sc = 2.0 / vv;
      DG_FOR(i, n) {
        double acc = (d0 - alpha) * Y[iq * ld + i];
        for (int j = 1; j < len; ++j) acc += Q.dv[iq + j] * Y[(iq + j) * ld + i];
        Q.zv[i] = acc * sc;
      }
      c.sync();
      DG_FOR(i, n) {
        const double wi = Q.zv[i];
        Y[iq * ld + i] -= (d0 - alpha) * wi;
        for (int j = 1; j < len; ++j) Y[(iq + j) * ld + i] -= Q.dv[iq + j] * wi;
      }
    }
    DG_FOR(i, iq) Rm[i * ld + iq] = Q.dv[i];
    if (c.tid() == 0) { Rm[iq * ld + iq] = alpha; Q.act[iq] = p; Q.is_act[p] = 1; }
    ++iq;
    c.sync();
  }
  if (iq == 0) return 0;
  game_G_times<SM>(c, D, E, Q.xq, Q.sl);                             // G x0
  while (iq > 0) {
    // r = g_W + G_W x0;  t = R^-T r (into zv);  lam = R^-1 t (into lam_act)      -- serial prototype on one thread
    if (c.tid() == 0) {
      for (int k = 0; k < iq; ++k) {
        double v = E.g[Q.act[k]] + Q.sl[Q.act[k]];
        for (int i = 0; i < k; ++i) v -= Rm[i * ld + k] * Q.zv[i];
        Q.zv[k] = v / Rm[k * ld + k];
      }
      for (int k = iq - 1; k >= 0; --k) {
        double v = Q.zv[k];
        for (int j = k + 1; j < iq; ++j) v -= Rm[k * ld + j] * Q.lam_act[j];
        Q.lam_act[k] = v / Rm[k * ld + k];
      }
    }
    c.sync();
    double worst = 0.0; int ldrop = -1;
    for (int k = 0; k < iq; ++k) { const double lk = Q.lam_act[k]; if (lk < worst) { worst = lk; ldrop = k; } }
    if (ldrop < 0) break;
    c.sync();
    // drop ldrop: same re-triangularisation as a partial step of the main loop
    if (c.tid() == 0) {
      Q.is_act[Q.act[ldrop]] = 0;
      for (int k = ldrop; k < iq - 1; ++k) Q.act[k] = Q.act[k + 1];
    }
    DG_FOR(i, iq) {
      for (int j = ldrop; j < iq - 1; ++j) Rm[i * ld + j] = Rm[i * ld + j + 1];
      Rm[i * ld + iq - 1] = 0.0;
    }
    c.sync();
    for (int j = ldrop; j < iq - 1; ++j) {
      const double a = Rm[j * ld + j], b = Rm[(j + 1) * ld + j];
      const double h = hypot(a, b);
      c.sync();
      if (h != 0.0) {
        const double cs = a / h, sn = b / h;
        for (int col = j + c.tid(); col < iq - 1; col += c.nt()) {
          const double r0 = Rm[j * ld + col], r1 = Rm[(j + 1) * ld + col];
          Rm[j * ld + col] = cs * r0 + sn * r1;
          Rm[(j + 1) * ld + col] = -sn * r0 + cs * r1;
        }
        DG_FOR(i, n) {
          const double j0 = Y[j * ld + i], j1 = Y[(j + 1) * ld + i];
          Y[j * ld + i] = cs * j0 + sn * j1;
          Y[(j + 1) * ld + i] = -sn * j0 + cs * j1;
        }
      }
      c.sync();
    }
    --iq;
  }
  // x = x0 + J1 t
  c.sync();
  DG_FOR(i, n) {
    double acc = 0.0;
    for (int j = 0; j < iq; ++j) acc += Y[j * ld + i] * Q.zv[j];
    Q.xq[i] += acc;
  }
  c.sync();
  return iq;
}
#define DG_WARM_PARAM , int warm = 0
#define DG_WARM_ARG(x) , (x)
#else
// (the experimental warm start is compiled out of the product: even unused, the extra call site cost 5 % in this function)
#define DG_WARM_PARAM
#define DG_WARM_ARG(x)
#endif

// H (symmetric positive definite) is expected in B.matA and is destroyed.
// returns 0 ok, 1 not PD, 2 infeasible, 3 iteration limit.  Output: Q.xq (du), Q.lam (l_hat).
template <bool SM>
DG_DEVN int qp_solve_gi(Cta& c, const Dims& D_, const EvalBuf& E_, const double* DG_RESTRICT qv,
                        const QpBuf& Q_, const LinBuf& B_, int* n_iter_out, int* n_active_out DG_WARM_PARAM) {
  // local copies: the tables live in shared memory and would otherwise be re-read after every store
  const EvalBuf E = E_; DG_SH_EVAL(E); const QpBuf Q = Q_; DG_SH_QP(Q); const LinBuf B = B_; DG_SH_LIN(B); const Dims D = D_;
  const int n = D.n, m = D.m, ld = B.ld;
  double* DG_RESTRICT Y = B.matB;
  double* DG_RESTRICT Rm = B.matA;
  if (!cholesky_lower<SM>(c, n, ld, B.matA, B.sp, B.part)) return 1;
  c.lap(PH_CHOL);
  tri_inverse<SM>(c, n, ld, B.matA, Y, B.sp, B.part);
  c.sync();
  // x = -J J' q = -Y' (Y q):   t = Y q (warp per row), x_i = -sum_j Y[j][i] t_j
  for (int j = c.warp(); j < n; j += c.nwarps()) {
    const double* DG_RESTRICT Yj = Y + j * ld;
    double p = 0.0;
    for (int i = c.lane(); i <= j; i += c.wsz) p += Yj[i] * qv[i];
    p = c.warp_sum(p);
    if (c.lane() == 0) Q.dv[j] = p;
  }
  DG_FOR(r, m) { Q.is_act[r] = 0; Q.lam[r] = 0.0; }
  c.sync();
  gi_cols_times<SM>(c, n, ld, Y, Q.dv, 0, B.part, Q.xq, -1.0);
  c.lap(PH_TRINV);
  int iq = 0, it = 0;
#ifdef DG_QP_WARM_START
  // active set of the previous QP of this instance: ids in Q.act[0..), count parked in Q.act[n] (see sqp_solve_*)
  if (warm) { const int nprev = Q.act[n]; c.sync(); if (nprev > 0 && nprev <= n) iq = gi_warm_start<SM>(c, D, E, Q, B, nprev); }
#endif
  const int max_iter = 10 * (n + m);
  const Split2 sp = split2(c, n);                 // one decomposition for every n-wide sweep of the loop (integer divisions)
  int status = 0;
  while (true) {
    // slacks of all constraints, most violated one
    game_G_times<SM>(c, D, E, Q.xq, Q.sl);
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
    game_G_row<SM>(c, D, E, p, Q.npv);                 // npv = G[p,:]  (normal is -npv)
    double lam_p = 0.0;
    bool added = false;
    while (!added) {
      if (++it > max_iter) { status = 3; break; }
      // d = J' n_p = -Y npv   (2D over the CTA: thread = row of Y, column groups interleave the columns);  also npv . x
      double dd_tail = 0.0, dd_all = 0.0, gx = 0.0;
      if constexpr (!SM) {
        // Y in the L2-resident workspace: the sweep is bound by L2 latency, so every lane keeps 16 loads in flight --
        // a warp takes four rows at a time, lanes along the rows (coalesced), 128 columns per pass, all loads issued
        // before the first use
        for (int j0 = 4 * c.warp(); j0 < n; j0 += 4 * c.nwarps()) {
          double acc[4] = {0.0, 0.0, 0.0, 0.0};
          for (int i0 = 0; i0 < n; i0 += 4 * c.wsz) {
            double y[4][4], v[4];
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                const int i = i0 + c.lane() + u * c.wsz;
                y[r][u] = (i < n && j0 + r < n) ? Y[(j0 + r) * ld + i] : 0.0;
              }
#pragma unroll
            for (int u = 0; u < 4; ++u) { const int i = i0 + c.lane() + u * c.wsz; v[u] = i < n ? Q.npv[i] : 0.0; }
#pragma unroll
            for (int r = 0; r < 4; ++r) acc[r] += (y[r][0] * v[0] + y[r][1] * v[1]) + (y[r][2] * v[2] + y[r][3] * v[3]);
          }
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            double a = c.warp_sum(acc[r]);
            if (c.lane() == 0 && j0 + r < n) {
              a = -a;
              Q.dv[j0 + r] = a;
              dd_all += a * a;
              if (j0 + r >= iq) dd_tail += a * a;
            }
          }
        }
      } else {
        for (int j = sp.i0; j < n; j += sp.istep) {
          const double* DG_RESTRICT Yj = Y + j * ld;
          double a0 = 0.0, a1 = 0.0;
          int i = sp.g;
          for (; i + sp.G < n; i += 2 * sp.G) { a0 += Yj[i] * Q.npv[i]; a1 += Yj[i + sp.G] * Q.npv[i + sp.G]; }
          if (i < n) a0 += Yj[i] * Q.npv[i];
          B.part[sp.g * sp.istep + j] = a0 + a1;
        }
        c.sync();
        if (sp.g == 0) {
          for (int j = sp.i0; j < n; j += sp.istep) {
            double acc = B.part[j];
            for (int g = 1; g < sp.G; ++g) acc += B.part[g * sp.istep + j];
            acc = -acc;
            Q.dv[j] = acc;
            dd_all += acc * acc;
            if (j >= iq) dd_tail += acc * acc;
          }
        }
      }
      DG_FOR(i, n) gx += Q.npv[i] * Q.xq[i];
      c.sum3(dd_tail, dd_all, gx);
      const double zn = dd_tail, dall = dd_all;
      // z = J[:, iq:] d[iq:]  (2D over the CTA);  r = R^-1 d[:iq]  (column-oriented back substitution on the last warp,
      // while the first column group combines the partial sums of z)
      {
        for (int i = sp.i0; i < n; i += sp.istep) {
          double a0 = 0.0, a1 = 0.0;
          int j = iq + sp.g;
          if constexpr (!SM) {
            // L2-resident Y: eight loads in flight per thread
            double a2 = 0.0, a3 = 0.0;
            for (; j + 7 * sp.G < n; j += 8 * sp.G) {
              double y[8];
#pragma unroll
              for (int u = 0; u < 8; ++u) y[u] = Y[(j + u * sp.G) * ld + i];
              a0 += y[0] * Q.dv[j] + y[4] * Q.dv[j + 4 * sp.G]; a1 += y[1] * Q.dv[j + sp.G] + y[5] * Q.dv[j + 5 * sp.G];
              a2 += y[2] * Q.dv[j + 2 * sp.G] + y[6] * Q.dv[j + 6 * sp.G]; a3 += y[3] * Q.dv[j + 3 * sp.G] + y[7] * Q.dv[j + 7 * sp.G];
            }
            for (; j + 3 * sp.G < n; j += 4 * sp.G) {
              const double y0 = Y[j * ld + i], y1 = Y[(j + sp.G) * ld + i], y2 = Y[(j + 2 * sp.G) * ld + i], y3 = Y[(j + 3 * sp.G) * ld + i];
              a0 += y0 * Q.dv[j]; a1 += y1 * Q.dv[j + sp.G]; a2 += y2 * Q.dv[j + 2 * sp.G]; a3 += y3 * Q.dv[j + 3 * sp.G];
            }
            a0 += a2; a1 += a3;
          }
          for (; j + sp.G < n; j += 2 * sp.G) { a0 += Y[j * ld + i] * Q.dv[j]; a1 += Y[(j + sp.G) * ld + i] * Q.dv[j + sp.G]; }
          if (j < n) a0 += Y[j * ld + i] * Q.dv[j];
          if (sp.G == 1) Q.zv[i] = a0 + a1;
          else B.part[sp.g * sp.istep + i] = a0 + a1;
        }
        if (sp.G > 1) {
          c.sync();
          if (sp.g == 0) {
            for (int i = sp.i0; i < n; i += sp.istep) {
              double acc = B.part[i];
              for (int g = 1; g < sp.G; ++g) acc += B.part[g * sp.istep + i];
              Q.zv[i] = acc;
            }
          }
        }
        if (c.warp() == c.nwarps() - 1) {
          // reciprocal pivots first (independent divisions), so that the serial chain below only multiplies;
          // they are parked in the tail of zv's partial-sum scratch (part[256..], beyond the G*cw <= 256 partial sums)
          double* DG_RESTRICT rinv = B.part + 256;
          for (int i = c.lane(); i < iq; i += c.wsz) { Q.rv[i] = Q.dv[i]; rinv[i] = 1.0 / Rm[i * ld + i]; }
          c.syncwarp();
          for (int i = iq - 1; i >= 0; --i) {
            const double ri = Q.rv[i] * rinv[i];
            c.syncwarp();
            if (c.lane() == 0) Q.rv[i] = ri;
            for (int j = c.lane(); j < i; j += c.wsz) Q.rv[j] -= Rm[j * ld + i] * ri;
            c.syncwarp();
          }
        }
      }
      c.sync();
      // dual step bound t1, primal step length t2
      double t1 = 1e300; int ldrop = -1;
      {
        double bv = 1e300; int bk = 0x7fffffff;
        DG_FOR(k, iq) {
          const double rk = Q.rv[k];
          if (rk > 0.0) { const double tk = Q.lam_act[k] / rk; if (tk < bv) { bv = tk; bk = k; } }
        }
        if (iq > 0) { c.argmin(bv, bk, t1, ldrop); if (!(t1 < 1e300)) ldrop = -1; }
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
          // w = (J2 v) * 2/(v'v) into npv (the entering row is not needed any more)
          DG_FOR(i, n) Q.npv[i] = (Q.zv[i] - alpha * Y[iq * ld + i]) * sc;
          c.sync();
          for (int i = sp.i0; i < n; i += sp.istep) {
            const double wi = Q.npv[i];
            double* DG_RESTRICT col = Y + iq * ld + i;
            int j = sp.g;
            if constexpr (!SM) {
              // L2-resident Y: eight independent read-modify-writes per step
              for (; j + 7 * sp.G < len; j += 8 * sp.G) {
                double y[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) y[u] = col[(j + u * sp.G) * ld];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                  const int jj = j + u * sp.G;
                  col[jj * ld] = y[u] - (jj == 0 ? d0 - alpha : Q.dv[iq + jj]) * wi;
                }
              }
            }
            for (; j < len; j += sp.G) col[j * ld] -= (j == 0 ? d0 - alpha : Q.dv[iq + j]) * wi;
          }
        }
        DG_FOR(i, iq) Rm[i * ld + iq] = Q.dv[i];
        if (c.tid() == 0) {
          Rm[iq * ld + iq] = alpha;
          Q.act[iq] = p; Q.is_act[p] = 1; Q.lam_act[iq] = lam_p;
        }
        ++iq;
        added = true;
        c.sync();
      } else {
        c.sync();
        // partial step: drop active constraint ldrop (Givens re-triangularisation), keep p
        if (c.tid() == 0) {
          Q.is_act[Q.act[ldrop]] = 0;
          for (int k = ldrop; k < iq - 1; ++k) { Q.act[k] = Q.act[k + 1]; Q.lam_act[k] = Q.lam_act[k + 1]; }
        }
        DG_FOR(i, iq) {
          for (int j = ldrop; j < iq - 1; ++j) Rm[i * ld + j] = Rm[i * ld + j + 1];
          Rm[i * ld + iq - 1] = 0.0;
        }
        c.sync();
        if (ldrop < iq - 1) {
          // Givens re-triangularisation.  The rotations form a serial chain only through R: one warp computes them and
          // rotates R (lanes along the columns), parking (cs, sn) in the partial-sum scratch; then every thread applies
          // the whole sequence to its own column of J (= column of the rows ldrop..iq-1 of Y) in registers -- two
          // barriers per drop instead of two per rotation.  Same arithmetic, same order as rotating row pairs one by one.
          double* DG_RESTRICT rot = B.part;
          if (c.warp() == 0) {
            for (int j = ldrop; j < iq - 1; ++j) {
              const double a = Rm[j * ld + j], b = Rm[(j + 1) * ld + j];
              const double h = hypot(a, b);
              double cs = 1.0, sn = 0.0;
              if (h != 0.0) { cs = a / h; sn = b / h; }
              c.syncwarp();                         // a, b read by every lane before the rows change
              if (h != 0.0) {
                for (int col = j + c.lane(); col < iq - 1; col += c.wsz) {
                  double r0 = Rm[j * ld + col], r1 = Rm[(j + 1) * ld + col];
                  Rm[j * ld + col] = cs * r0 + sn * r1;
                  Rm[(j + 1) * ld + col] = -sn * r0 + cs * r1;
                }
              }
              if (c.lane() == 0) { rot[2 * (j - ldrop)] = cs; rot[2 * (j - ldrop) + 1] = sn; }
              c.syncwarp();
            }
          }
          c.sync();
          DG_FOR(i, n) {
            double t = Y[ldrop * ld + i];
            for (int j = ldrop; j < iq - 1; ++j) {
              const double cs = rot[2 * (j - ldrop)], sn = rot[2 * (j - ldrop) + 1];
              const double u1 = Y[(j + 1) * ld + i];
              if (cs == 1.0 && sn == 0.0) { Y[j * ld + i] = t; t = u1; }      // h == 0: no rotation
              else { Y[j * ld + i] = cs * t + sn * u1; t = -sn * t + cs * u1; }
            }
            Y[(iq - 1) * ld + i] = t;
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
#ifdef DG_QP_WARM_START
  if (c.tid() == 0) Q.act[n] = status == 0 ? iq : 0;               // remembered for a warm start of the next QP
#endif
  if (n_iter_out) *n_iter_out = it;
  if (n_active_out) *n_active_out = iq;
  c.lap(PH_GI);
  return status;
}
