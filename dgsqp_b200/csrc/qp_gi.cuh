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

// ---- warm start of the dual active-set method from the active set of the previous QP of this instance ------------------
// (SolverParams::qp_warm; BASELINE north_star: "per-instance warm start").  Successive QPs of one instance -- the next SQP
// iteration, the relaxed steps of the watchdog -- share most of their active set, and the cold method spends one
// O(n^2) iteration per constraint it adds.  Starting from the unconstrained minimiser x0 in Q.xq and Y = L^-1:
//   (1) D = J' N_W = -Y G_W' for all previous constraints at once (one product per column, no step-length logic);
//   (2) Householder QR of D (n x k, in matA, right-looking; a column that is linearly dependent on the accepted ones --
//       same test as the main loop -- is skipped): R is the factor the main loop continues with, the reflectors
//       below the diagonal are applied to Y in ONE barrier-free pass, thread = column of Y;
//   (3) the equality-constrained minimiser on that set in the cancellation-free form of gi_polish:
//       c = J'q,  t = R^-T g_W,  lam = R^-1 (t + c1),  x = J1 t - J2 c2;
//   (4) while a multiplier is negative: drop the most negative one (Givens re-triangularisation, c rotated alike) and
//       redo (3).
// (x, lam >= 0, W) is then an S-pair, from which the method continues unchanged.  The QP is strictly convex, so the
// solution is the one the cold start reaches (up to rounding); only the path differs.
// Scratch: previous ids in Q.rv (as ints), reflector scalars v0 / 2/(v'v) in Q.dv / Q.zv (then c / [t; -c2]), g_W in
// Q.npv; the entering row of (1) goes through Q.lam, which is only written after the main loop.
// Returns the number of active constraints; Q.act / Q.is_act / Q.lam_act / Q.xq are set accordingly.
template <bool SM>
DG_DEVN int gi_warm_start(Cta& c, const Dims& D_, const EvalBuf& E_, const double* DG_RESTRICT qv, const QpBuf& Q_, const LinBuf& B_, int nprev) {
  const EvalBuf E = E_; DG_SH_EVAL(E); const QpBuf Q = Q_; DG_SH_QP(Q); const LinBuf B = B_; DG_SH_LIN(B); const Dims D = D_;
  const int n = D.n, ld = B.ld;
  double* DG_RESTRICT Y = B.matB;
  double* DG_RESTRICT Rm = B.matA;
  int* DG_RESTRICT ids = (int*)Q.rv;
  double* DG_RESTRICT v0s = Q.dv;
  double* DG_RESTRICT scs = Q.zv;
  double* DG_RESTRICT row = Q.lam;                 // m >= n doubles, free until the end of the QP
  double* DG_RESTRICT rt = Q.npv;
  c.sync();
  DG_FOR(k, nprev) ids[k] = Q.act[k];
  c.sync();
  // (1) D[:, k] = -Y G[p_k, :]'  into Rm[j*ld + k];  rt[k] = g_p
  const Split2 sp = split2(c, n);
  for (int k = 0; k < nprev; ++k) {
    const int p = ids[k];
    game_G_row<SM>(c, D, E, p, row);
    if constexpr (!SM) {
      // Y may live in the L2-resident workspace: warp per row, lanes along the row (coalesced)
      for (int j = c.warp(); j < n; j += c.nwarps()) {
        double acc = 0.0;
        for (int i = c.lane(); i < n; i += c.wsz) acc += Y[j * ld + i] * row[i];
        acc = c.warp_sum(acc);
        if (c.lane() == 0) Rm[j * ld + k] = -acc;
      }
    } else {
      for (int j = sp.i0; j < n; j += sp.istep) {
        const double* DG_RESTRICT Yj = Y + j * ld;
        double a0 = 0.0, a1 = 0.0;
        int i = sp.g;
        for (; i + sp.G < n; i += 2 * sp.G) { a0 += Yj[i] * row[i]; a1 += Yj[i + sp.G] * row[i + sp.G]; }
        if (i < n) a0 += Yj[i] * row[i];
        if (sp.G == 1) Rm[j * ld + k] = -(a0 + a1);
        else B.part[sp.g * sp.istep + j] = a0 + a1;
      }
      if (sp.G > 1) {
        c.sync();
        if (sp.g == 0) {
          for (int j = sp.i0; j < n; j += sp.istep) {
            double acc = B.part[j];
            for (int g = 1; g < sp.G; ++g) acc += B.part[g * sp.istep + j];
            Rm[j * ld + k] = -acc;
          }
        }
      }
    }
    if (c.tid() == 0) rt[k] = E.g[p];
    c.sync();                                      // the next row overwrites `row` / `part`
  }
  // (2) Householder QR with the dependence test of the main loop; accepted column iq <- source column k
  int iq = 0;
  for (int k = 0; k < nprev; ++k) {
    double zn = 0.0, dall = 0.0;
    DG_FOR(j, n) { const double dj = Rm[j * ld + k]; dall += dj * dj; if (j >= iq) zn += dj * dj; }
    c.sum2(zn, dall);
    if (!(zn > DG_QP_DEP_TOL * dall && zn > 0.0)) continue;          // linearly dependent on the accepted ones
    const double d0 = Rm[iq * ld + k];
    double alpha = sqrt(zn);
    if (d0 > 0.0) alpha = -alpha;
    const double v0 = d0 - alpha;
    const double vv = 2.0 * (zn - alpha * d0);
    const double sc = vv > 0.0 ? 2.0 / vv : 0.0;
    c.sync();                                                       // d0 read by everyone before the column moves
    // apply (I - sc v v') to the remaining source columns: warp per column, lanes along the rows
    for (int k2 = k + 1 + c.warp(); k2 < nprev; k2 += c.nwarps()) {
      double acc = 0.0;
      for (int j = iq + c.lane(); j < n; j += c.wsz) acc += (j == iq ? v0 : Rm[j * ld + k]) * Rm[j * ld + k2];
      acc = c.warp_sum(acc) * sc;
      for (int j = iq + c.lane(); j < n; j += c.wsz) Rm[j * ld + k2] -= (j == iq ? v0 : Rm[j * ld + k]) * acc;
    }
    c.sync();
    // move the column into place: rows < iq of R, alpha on the diagonal, the reflector below it
    if (iq != k) { DG_FOR(j, n) if (j != iq) Rm[j * ld + iq] = Rm[j * ld + k]; }
    if (c.tid() == 0) {
      Rm[iq * ld + iq] = alpha; v0s[iq] = v0; scs[iq] = sc;
      Q.act[iq] = ids[k]; Q.is_act[ids[k]] = 1; rt[iq] = rt[k];
    }
    ++iq;
    c.sync();
  }
  if (iq == 0) return 0;
  // Y <- P_{iq-1} .. P_0 Y: reflector k acts on rows k..n-1; the columns of Y are independent -> thread per column,
  // no barrier.  Column i of Y = column i of the rows of J', i.e. J[:, k:] <- J[:, k:] P_k like a full step of the loop.
  DG_FOR(i, n) {
    for (int k = 0; k < iq; ++k) {
      const double v0 = v0s[k];
      double a0 = v0 * Y[k * ld + i], a1 = 0.0;
      int j = k + 1;
      for (; j + 1 < n; j += 2) { a0 += Rm[j * ld + k] * Y[j * ld + i]; a1 += Rm[(j + 1) * ld + k] * Y[(j + 1) * ld + i]; }
      if (j < n) a0 += Rm[j * ld + k] * Y[j * ld + i];
      const double w = (a0 + a1) * scs[k];
      Y[k * ld + i] -= v0 * w;
      for (j = k + 1; j < n; ++j) Y[j * ld + i] -= Rm[j * ld + k] * w;
    }
  }
  c.sync();
  // c = J'q = Y q into Q.dv (the reflector scalars are dead now)
  double* DG_RESTRICT cv = Q.dv;
  if constexpr (!SM) {
    for (int j = c.warp(); j < n; j += c.nwarps()) {
      double acc = 0.0;
      for (int i = c.lane(); i < n; i += c.wsz) acc += Y[j * ld + i] * qv[i];
      acc = c.warp_sum(acc);
      if (c.lane() == 0) cv[j] = acc;
    }
  } else {
    for (int j = sp.i0; j < n; j += sp.istep) {
      const double* DG_RESTRICT Yj = Y + j * ld;
      double a0 = 0.0, a1 = 0.0;
      int i = sp.g;
      for (; i + sp.G < n; i += 2 * sp.G) { a0 += Yj[i] * qv[i]; a1 += Yj[i + sp.G] * qv[i + sp.G]; }
      if (i < n) a0 += Yj[i] * qv[i];
      B.part[sp.g * sp.istep + j] = a0 + a1;
    }
    c.sync();
    if (sp.g == 0) {
      for (int j = sp.i0; j < n; j += sp.istep) {
        double acc = B.part[j];
        for (int g = 1; g < sp.G; ++g) acc += B.part[g * sp.istep + j];
        cv[j] = acc;
      }
    }
  }
  c.sync();
  // (3)/(4) multipliers on the accepted set; drop negative ones
  double* DG_RESTRICT tv = Q.zv;                   // t = R^-T g_W
  double* DG_RESTRICT rinv = B.part + 256;         // 1 / R_kk  (part holds max(512, 2n) doubles; the rotations use part[0..2n))
  while (true) {
    if (c.warp() == 0) {
      for (int k = c.lane(); k < iq; k += c.wsz) { tv[k] = rt[k]; rinv[k] = 1.0 / Rm[k * ld + k]; }
      c.syncwarp();
      // forward substitution with R' (right-looking: row k of R is contiguous)
      for (int k = 0; k < iq; ++k) {
        const double tk = tv[k] * rinv[k];
        c.syncwarp();
        if (c.lane() == 0) tv[k] = tk;
        for (int j = k + 1 + c.lane(); j < iq; j += c.wsz) tv[j] -= Rm[k * ld + j] * tk;
        c.syncwarp();
      }
      // back substitution lam = R^-1 (t + c1) (column-oriented, like the main loop)
      for (int k = c.lane(); k < iq; k += c.wsz) Q.lam_act[k] = tv[k] + cv[k];
      c.syncwarp();
      for (int k = iq - 1; k >= 0; --k) {
        const double lk = Q.lam_act[k] * rinv[k];
        c.syncwarp();
        if (c.lane() == 0) Q.lam_act[k] = lk;
        for (int j = c.lane(); j < k; j += c.wsz) Q.lam_act[j] -= Rm[j * ld + k] * lk;
        c.syncwarp();
      }
    }
    c.sync();
    double worst; int ldrop;
    {
      double bv = 0.0; int bk = 0x7fffffff;
      DG_FOR(k, iq) { const double lk = Q.lam_act[k]; if (lk < bv) { bv = lk; bk = k; } }
      c.argmin(bv, bk, worst, ldrop);
    }
    if (!(worst < 0.0)) break;
    // drop ldrop: shift the columns of R, re-triangularise with Givens rotations, rotate the rows of Y (and c) alike
    if (c.tid() == 0) {
      Q.is_act[Q.act[ldrop]] = 0;
      for (int k = ldrop; k < iq - 1; ++k) { Q.act[k] = Q.act[k + 1]; rt[k] = rt[k + 1]; }
    }
    DG_FOR(i, iq) {
      for (int j = ldrop; j < iq - 1; ++j) Rm[i * ld + j] = Rm[i * ld + j + 1];
      Rm[i * ld + iq - 1] = 0.0;
    }
    c.sync();
    if (ldrop < iq - 1) {
      double* DG_RESTRICT rot = B.part;
      if (c.warp() == 0) {
        for (int j = ldrop; j < iq - 1; ++j) {
          const double a = Rm[j * ld + j], b = Rm[(j + 1) * ld + j];
          const double h = hypot(a, b);
          double cs = 1.0, sn = 0.0;
          if (h != 0.0) { cs = a / h; sn = b / h; }
          c.syncwarp();
          if (h != 0.0) {
            for (int col = j + c.lane(); col < iq - 1; col += c.wsz) {
              double r0 = Rm[j * ld + col], r1 = Rm[(j + 1) * ld + col];
              Rm[j * ld + col] = cs * r0 + sn * r1;
              Rm[(j + 1) * ld + col] = -sn * r0 + cs * r1;
            }
          }
          if (c.lane() == 0) {
            rot[2 * (j - ldrop)] = cs; rot[2 * (j - ldrop) + 1] = sn;
            const double c0 = cv[j], c1 = cv[j + 1];
            cv[j] = cs * c0 + sn * c1; cv[j + 1] = -sn * c0 + cs * c1;
          }
          c.syncwarp();
        }
      }
      c.sync();
      DG_FOR(i, n) {
        double t = Y[ldrop * ld + i];
        for (int j = ldrop; j < iq - 1; ++j) {
          const double cs = rot[2 * (j - ldrop)], sn = rot[2 * (j - ldrop) + 1];
          const double u1 = Y[(j + 1) * ld + i];
          if (cs == 1.0 && sn == 0.0) { Y[j * ld + i] = t; t = u1; }
          else { Y[j * ld + i] = cs * t + sn * u1; t = -sn * t + cs * u1; }
        }
        Y[(iq - 1) * ld + i] = t;
      }
      c.sync();
    }
    --iq;
    if (iq == 0) break;
  }
  // x = J1 t - J2 c2 = Y' [t; -c2]
  c.sync();
  for (int j = iq + c.tid(); j < n; j += c.nt()) tv[j] = -cv[j];
  c.sync();
  gi_cols_times<SM>(c, n, ld, Y, tv, 0, B.part, Q.xq, 1.0);
  return iq;
}

// Cholesky of H (in B.matA, destroyed), Y = L^-1 into B.matB, unconstrained minimiser into Q.xq, empty active set.
// Returns false (uniformly) when H is not positive definite.
template <bool SM>
DG_DEVN bool qp_factor(Cta& c, const Dims& D_, const double* DG_RESTRICT qv, const QpBuf& Q_, const LinBuf& B_) {
  // local copies: the tables live in shared memory and would otherwise be re-read after every store
  const QpBuf Q = Q_; DG_SH_QP(Q); const LinBuf B = B_; DG_SH_LIN(B); const Dims D = D_;
  const int n = D.n, m = D.m, ld = B.ld;
  double* DG_RESTRICT Y = B.matB;
  {
    bool ok = false;
#ifndef DG_HOSTSIM
    if (!cholesky_regs_dispatch<SM>(c, n, B, ok))
#endif
    ok = cholesky_lower<SM>(c, n, ld, B.matA, B.sp, B.part);
    if (!ok) return false;
  }
  c.lap(PH_CHOL);
  tri_inverse<SM>(c, n, ld, B.matA, Y, B.sp, B.part);
  c.sync();
  c.lapf(PH_TRINV);
  // x = -J J' q = -Y' (Y q):   t = Y q (warp per row), x_i = -sum_j Y[j][i] t_j
  for (int j = c.warp(); j < n; j += c.nwarps()) {
    const double* DG_RESTRICT Yj = Y + j * ld;
    double p = 0.0;
    for (int i = c.lane(); i <= j; i += c.wsz) p += Yj[i] * qv[i];
    p = c.warp_sum(p);
    if (c.lane() == 0) Q.dv[j] = p;
  }
  DG_FOR(r, m) Q.is_act[r] = 0;
  c.sync();
  gi_cols_times<SM>(c, n, ld, Y, Q.dv, 0, B.part, Q.xq, -1.0);
  c.lap2(PH_TRINV, PH_QP_X0);
  return true;
}

// The dual active-set iteration from the S-pair (Q.xq, Q.act[0..iq0), Q.lam_act) with factors Y = J' (B.matB) and R (B.matA).
// returns 0 ok, 2 infeasible, 3 iteration limit.  Output: Q.xq (du), Q.lam (l_hat).
template <bool SM>
DG_DEVN int qp_gi_loop(Cta& c, const Dims& D_, const EvalBuf& E_, const QpBuf& Q_, const LinBuf& B_, int iq0,
                       int* n_iter_out, int* n_active_out) {
  // local copies: the tables live in shared memory and would otherwise be re-read after every store
  const EvalBuf E = E_; DG_SH_EVAL(E); const QpBuf Q = Q_; DG_SH_QP(Q); const LinBuf B = B_; DG_SH_LIN(B); const Dims D = D_;
  const int n = D.n, m = D.m, ld = B.ld;
  double* DG_RESTRICT Y = B.matB;
  double* DG_RESTRICT Rm = B.matA;
  DG_FOR(r, m) Q.lam[r] = 0.0;
  int iq = iq0, it = 0;
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
    c.lapf(PH_GI_SLACK);
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
      c.lapf(PH_GI_DZ);
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
      c.lapf(PH_GI_STEP);
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
        c.lapf(PH_GI_ADD);
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
        c.lapf(PH_GI_DROP);
      }
    }
    if (status != 0) break;
  }
  if (status == 0) {
    c.sync();
    DG_FOR(k, iq) Q.lam[Q.act[k]] = Q.lam_act[k];
    c.sync();
  }
  if (c.tid() == 0) Q.act[n] = status == 0 ? iq : 0;               // remembered for a warm start of the next QP
  if (n_iter_out) *n_iter_out = it;
  if (n_active_out) *n_active_out = iq;
  c.lap(PH_GI);
  return status;
}

// Polish: the KKT point of the final active set W re-evaluated from the final factors without the cancellation the
// iteration carries.  The method starts at the unconstrained minimiser x0 = -H^-1 q and moves by x += t z; when H is
// ill-conditioned (reg = 0 games: eigenvalues at the 1e-10 floor, condition 1e11) x0 is ~1e10 |q| along the near-null
// directions and the steps cancel it, which leaves errors of 1e-5..1e-3 in x and 1e-2 in lam.  With J = [J1 J2], J'N = [R; 0]
// (N = -G_W') and c = J'q the same point is
//     x = J1 R^-T g_W - J2 c2,        lam_W = R^-1 (R^-T g_W + c1)
// in which the large components c1 never meet (measured against an extended-precision KKT solve on the merge game: 1e-10
// in x, 1e-8 in lam).  This is the role OSQP's polish step has in the reference (DGSQP.py:186, polish=True: reduced KKT
// system of the active set + iterative refinement); the oracle applies the same formula (oracle/qp.py).
template <bool SM>
DG_DEVN void gi_polish(Cta& c, const Dims& D_, const EvalBuf& E_, const double* DG_RESTRICT qv, const QpBuf& Q_, const LinBuf& B_, int iq) {
  const EvalBuf E = E_; DG_SH_EVAL(E); const QpBuf Q = Q_; DG_SH_QP(Q); const LinBuf B = B_; DG_SH_LIN(B); const Dims D = D_;
  const int n = D.n, ld = B.ld;
  double* DG_RESTRICT Y = B.matB;
  double* DG_RESTRICT Rm = B.matA;
  c.sync();
  // c = J' q = Y q into Q.dv
  if constexpr (!SM) {
    for (int j = c.warp(); j < n; j += c.nwarps()) {
      double acc = 0.0;
      for (int i = c.lane(); i < n; i += c.wsz) acc += Y[j * ld + i] * qv[i];
      acc = c.warp_sum(acc);
      if (c.lane() == 0) Q.dv[j] = acc;
    }
    c.sync();
  } else {
    const Split2 sp = split2(c, n);
    for (int j = sp.i0; j < n; j += sp.istep) {
      const double* DG_RESTRICT Yj = Y + j * ld;
      double a0 = 0.0, a1 = 0.0;
      int i = sp.g;
      for (; i + sp.G < n; i += 2 * sp.G) { a0 += Yj[i] * qv[i]; a1 += Yj[i + sp.G] * qv[i + sp.G]; }
      if (i < n) a0 += Yj[i] * qv[i];
      B.part[sp.g * sp.istep + j] = a0 + a1;
    }
    c.sync();
    if (sp.g == 0) {
      for (int j = sp.i0; j < n; j += sp.istep) {
        double acc = B.part[j];
        for (int g = 1; g < sp.G; ++g) acc += B.part[g * sp.istep + j];
        Q.dv[j] = acc;
      }
    }
    c.sync();
  }
  // w = [R^-T g_W ; -c2] into Q.zv,  lam_W = R^-1 (w1 + c1) into Q.lam_act
  if (c.warp() == 0) {
    double* DG_RESTRICT tv = Q.zv;
    double* DG_RESTRICT rinv = B.part + 256;
    for (int k = c.lane(); k < iq; k += c.wsz) { tv[k] = E.g[Q.act[k]]; rinv[k] = 1.0 / Rm[k * ld + k]; }
    c.syncwarp();
    for (int k = 0; k < iq; ++k) {
      const double tk = tv[k] * rinv[k];
      c.syncwarp();
      if (c.lane() == 0) tv[k] = tk;
      for (int j = k + 1 + c.lane(); j < iq; j += c.wsz) tv[j] -= Rm[k * ld + j] * tk;
      c.syncwarp();
    }
    for (int k = c.lane(); k < iq; k += c.wsz) Q.lam_act[k] = tv[k] + Q.dv[k];
    c.syncwarp();
    for (int k = iq - 1; k >= 0; --k) {
      const double lk = Q.lam_act[k] * rinv[k];
      c.syncwarp();
      if (c.lane() == 0) Q.lam_act[k] = lk;
      for (int j = c.lane(); j < k; j += c.wsz) Q.lam_act[j] -= Rm[j * ld + k] * lk;
      c.syncwarp();
    }
  } else {
    for (int j = iq + c.tid() - c.wsz; j < n; j += c.nt() - c.wsz) Q.zv[j] = -Q.dv[j];
  }
  if (c.nt() <= c.wsz) { for (int j = iq; j < n; ++j) Q.zv[j] = -Q.dv[j]; }     // one-warp (and host) builds: warp 0 does both
  c.sync();
  gi_cols_times<SM>(c, n, ld, Y, Q.zv, 0, B.part, Q.xq, 1.0);
  DG_FOR(k, iq) Q.lam[Q.act[k]] = Q.lam_act[k];
  c.sync();
}

// _solve_qp's solver call.  H (symmetric positive definite) is expected in B.matA and is destroyed.
// warm != 0: start from the active set of the previous QP of this instance (ids in Q.act[0..), count parked in Q.act[n];
// the SQP drivers reset the count at the start of an instance).
// returns 0 ok, 1 not PD, 2 infeasible, 3 iteration limit.  Output: Q.xq (du), Q.lam (l_hat).
template <bool SM>
DG_DEV int qp_solve_gi(Cta& c, const Dims& D, const EvalBuf& E, const double* qv, const QpBuf& Q, const LinBuf& B,
                       int* n_iter_out, int* n_active_out, int warm) {
  if (!qp_factor<SM>(c, D, qv, Q, B)) return 1;
  int iq = 0;
  if (warm) {
    const int nprev = Q.act[D.n];
    if (nprev > 0 && nprev <= D.n) { iq = gi_warm_start<SM>(c, D, E, qv, Q, B, nprev); c.lapf(PH_QP_WARM); }
  }
  int na = 0;
  const int st = qp_gi_loop<SM>(c, D, E, Q, B, iq, n_iter_out, &na);
  if (n_active_out) *n_active_out = na;
  if (st == 0) { gi_polish<SM>(c, D, E, qv, Q, B, na); c.lap(PH_GI); }
  return st;
}
