// CTA-cooperative dense FP64 kernels for one n x n problem (n ~ 100-200), row-major storage.
//
// nearest_pd() replaces DGSQP._nearestPD (DGSQP/solvers/DGSQP.py:1290-1296: B=(A+A')/2; eigh;
// negative eigenvalues -> floor; U diag(s) U'; re-symmetrise).  Only the negative part of the
// spectrum changes, so   nearestPD(A) = B + sum_{s_i<0} (floor - s_i) u_i u_i'   and only the
// negative eigenpairs are needed: Householder tridiagonalisation, Sturm-count multisection for the
// negative eigenvalues, inverse iteration on the tridiagonal matrix, back-transformation.  When the
// Sturm count at 0 is zero the projection is the identity and the whole eigen-solve is skipped.
#pragma once
#include "cta.cuh"

#define DG_EIG_CHUNK 16     // eigenvectors computed concurrently by inverse iteration

struct LinBuf {
  double* W;      // n*n   tridiagonalisation workspace (holds the reflectors afterwards)
  double* dg;     // n     tridiagonal diagonal
  double* od;     // n     off-diagonal (od[k] couples k, k+1)
  double* tau;    // n
  double* pv;     // n     scratch vector
  double* wv;     // n     scratch vector
  double* lam;    // n     negative eigenvalues (ascending)
  double* Z;      // DG_EIG_CHUNK*n  eigenvectors of the chunk
  double* itw;    // DG_EIG_CHUNK*5*n  inverse-iteration factor storage
};

// Householder reduction of the symmetric matrix W (full storage, both triangles kept consistent)
// to tridiagonal form; reflector k is stored in W[k+2.., k] (v[0] = 1 implicit) with tau[k].
DG_DEVN void sym_tridiag(Cta& c, int n, const LinBuf& B) {
  double* W = B.W;
  for (int k = 0; k + 1 < n; ++k) {
    const int len = n - k - 1;            // x = W[k+1.., k]
    double part = 0.0;
    for (int i = c.tid + 1; i < len; i += c.nt) { double xv = W[(k + 1 + i) * n + k]; part += xv * xv; }
    double xn2 = c.sum(part);
    double alpha = W[(k + 1) * n + k];
    double tauk = 0.0, beta = alpha, scale = 0.0;
    if (xn2 > 0.0) {
      beta = -copysign(sqrt(alpha * alpha + xn2), alpha);
      tauk = (beta - alpha) / beta;
      scale = 1.0 / (alpha - beta);
    }
    // v into pv (v[0]=1), also stored back into column k
    DG_FOR(i, len) {
      double vi = i == 0 ? 1.0 : W[(k + 1 + i) * n + k] * scale;
      B.pv[i] = vi;
      if (i > 0) W[(k + 1 + i) * n + k] = vi;
    }
    if (c.tid == 0) { B.dg[k] = W[k * n + k]; B.od[k] = beta; B.tau[k] = tauk; }
    c.sync();
    if (tauk != 0.0) {
      // p = tau * A22 v  (column access: thread i reads W[j][i], coalesced across threads)
      double pdot = 0.0;
      DG_FOR(i, len) {
        double acc = 0.0;
        for (int j = 0; j < len; ++j) acc += W[(k + 1 + j) * n + (k + 1 + i)] * B.pv[j];
        acc *= tauk;
        B.wv[i] = acc;
        pdot += acc * B.pv[i];
      }
      double pv_dot = c.sum(pdot);
      double hk = 0.5 * tauk * pv_dot;
      DG_FOR(i, len) B.wv[i] -= hk * B.pv[i];
      c.sync();
      // A22 -= v w' + w v'   (thread per column)
      DG_FOR(j, len) {
        double vj = B.pv[j], wj = B.wv[j];
        for (int i = 0; i < len; ++i) W[(k + 1 + i) * n + (k + 1 + j)] -= B.pv[i] * wj + B.wv[i] * vj;
      }
      c.sync();
    }
  }
  if (c.tid == 0) { B.dg[n - 1] = W[(n - 1) * n + (n - 1)]; B.od[n - 1] = 0.0; }
  c.sync();
}

// number of eigenvalues of tridiag(dg, od) that are < x
DG_DEV int sturm_count(int n, const double* dg, const double* od, double x, double pivmin) {
  int cnt = 0;
  double q = dg[0] - x;
  if (fabs(q) < pivmin) q = -pivmin;
  if (q < 0.0) ++cnt;
  for (int i = 1; i < n; ++i) {
    q = dg[i] - x - od[i - 1] * od[i - 1] / q;
    if (fabs(q) < pivmin) q = -pivmin;
    if (q < 0.0) ++cnt;
  }
  return cnt;
}

// Solve (T - lam I) y = b in place (y overwrites b) by Gaussian elimination with partial pivoting.
// fw: 5*n scratch (diag, sup1, sup2, mult, swapped)
DG_DEV void tridiag_shift_solve(int n, const double* dg, const double* od, double lam, double tiny,
                                double* fw, double* y, bool refactor) {
  double* a = fw; double* b1 = fw + n; double* b2 = fw + 2 * n; double* ml = fw + 3 * n; double* sw = fw + 4 * n;
  if (refactor) {
    // rows: row i = [sub_i, diag_i, sup_i];  work row "cur" = (p, q, r) starting at column i
    double p = dg[0] - lam, q = n > 1 ? od[0] : 0.0, r = 0.0;
    for (int i = 0; i + 1 < n; ++i) {
      double sub = od[i], nd = dg[i + 1] - lam, ns = i + 2 < n ? od[i + 1] : 0.0;
      if (fabs(sub) > fabs(p)) {
        // swap: pivot row is the next row (sub, nd, ns)
        double m = p / sub;
        a[i] = sub; b1[i] = nd; b2[i] = ns; ml[i] = m; sw[i] = 1.0;
        p = q - m * nd; q = r - m * ns; r = 0.0;
      } else {
        if (p == 0.0) p = tiny;
        double m = sub / p;
        a[i] = p; b1[i] = q; b2[i] = r; ml[i] = m; sw[i] = 0.0;
        p = nd - m * q; q = ns - m * r; r = 0.0;
      }
    }
    if (fabs(p) < tiny) p = p < 0.0 ? -tiny : tiny;
    a[n - 1] = p; b1[n - 1] = 0.0; b2[n - 1] = 0.0;
  }
  // forward
  for (int i = 0; i + 1 < n; ++i) {
    if (sw[i] != 0.0) { double t = y[i]; y[i] = y[i + 1]; y[i + 1] = t - ml[i] * y[i]; }
    else y[i + 1] -= ml[i] * y[i];
  }
  // backward
  for (int i = n - 1; i >= 0; --i) {
    double t = y[i];
    if (i + 1 < n) t -= b1[i] * y[i + 1];
    if (i + 2 < n) t -= b2[i] * y[i + 2];
    double piv = a[i];
    if (fabs(piv) < tiny) piv = piv < 0.0 ? -tiny : tiny;
    y[i] = t / piv;
  }
}

// Hm <- nearestPD(Qraw) + reg*I.  Returns the number of negative eigenvalues (uniform across threads).
DG_DEVN int nearest_pd(Cta& c, int n, const double* Qraw, double* Hm, const LinBuf& B, double floor_val,
                       double reg, bool conv_approx) {
  // symmetric part into W and Hm
  DG_FOR(t, n * n) {
    int i = t / n, j = t - i * n;
    double sv = 0.5 * (Qraw[i * n + j] + Qraw[j * n + i]);
    B.W[t] = sv; Hm[t] = sv;
  }
  c.sync();
  int nneg = 0;
  if (conv_approx) {
    sym_tridiag(c, n, B);
    // norms / pivmin
    double tn = 0.0;
    DG_FOR(i, n) {
      double r = fabs(B.dg[i]) + fabs(B.od[i]) + (i > 0 ? fabs(B.od[i - 1]) : 0.0);
      tn = fmax(tn, r);
    }
    const double tnorm = c.max(tn);
    const double pivmin = 2.2250738585072014e-308 * fmax(1.0, tnorm * tnorm);
    nneg = sturm_count(n, B.dg, B.od, 0.0, pivmin);      // every thread computes the same count
    if (nneg > 0) {
      // --- negative eigenvalues by multisection on Sturm counts
      int rounds = (int)ceil(56.0 * 0.6931471805599453 / log((double)c.nt + 1.0));
      if (rounds < 1) rounds = 1;
      double lo_all = -tnorm * 1.0000001 - pivmin;
      for (int j = 0; j < nneg; ++j) {
        double a = j == 0 ? lo_all : B.lam[j - 1], b = 0.0;
        for (int rd = 0; rd < rounds; ++rd) {
          double h = (b - a) / (double)(c.nt + 1);
          double xt = a + h * (double)(c.tid + 1);
          int ct = sturm_count(n, B.dg, B.od, xt, pivmin);
          int first = c.imin(ct >= j + 1 ? c.tid : c.nt);
          double na = first > 0 ? a + h * (double)first : a;              // x_{first-1}
          double nb = first < c.nt ? a + h * (double)(first + 1) : b;     // x_first
          a = na; b = nb;
          if (!(b - a > 4.4e-16 * fmax(fabs(a), fabs(b)))) break;
        }
        c.sync();
        if (c.tid == 0) B.lam[j] = 0.5 * (a + b);
        c.sync();
      }
      // --- eigenvectors in chunks: inverse iteration (thread per vector), MGS inside clusters,
      //     back-transformation, rank-one corrections of Hm
      const double tiny = fmax(tnorm, 1.0) * 1.1e-16;
      for (int j0 = 0; j0 < nneg; j0 += DG_EIG_CHUNK) {
        int kc = nneg - j0 < DG_EIG_CHUNK ? nneg - j0 : DG_EIG_CHUNK;
        for (int itn = 0; itn < 4; ++itn) {
          DG_FOR(jj, kc) {
            double* z = B.Z + jj * n;
            if (itn == 0)
              for (int i = 0; i < n; ++i) z[i] = 1.0 + 0.37 * (double)(((i + 1) * 7919 + (j0 + jj) * 104729) % 97) / 97.0;
            tridiag_shift_solve(n, B.dg, B.od, B.lam[j0 + jj], tiny, B.itw + jj * 5 * n, z, itn == 0);
            double nr = 0.0;
            for (int i = 0; i < n; ++i) nr += z[i] * z[i];
            nr = 1.0 / sqrt(nr);
            for (int i = 0; i < n; ++i) z[i] *= nr;
          }
          c.sync();
          // modified Gram-Schmidt against earlier vectors of the same cluster (same chunk)
          if (kc > 1) {
            if (c.tid == 0) {
              for (int jj = 1; jj < kc; ++jj) {
                double* z = B.Z + jj * n;
                bool changed = false;
                for (int ii = 0; ii < jj; ++ii) {
                  if (fabs(B.lam[j0 + jj] - B.lam[j0 + ii]) > 1e-3 * tnorm) continue;
                  const double* zi = B.Z + ii * n;
                  double dt = 0.0;
                  for (int i = 0; i < n; ++i) dt += zi[i] * z[i];
                  for (int i = 0; i < n; ++i) z[i] -= dt * zi[i];
                  changed = true;
                }
                if (changed) {
                  double nr = 0.0;
                  for (int i = 0; i < n; ++i) nr += z[i] * z[i];
                  nr = 1.0 / sqrt(nr);
                  for (int i = 0; i < n; ++i) z[i] *= nr;
                }
              }
            }
            c.sync();
          }
        }
        // back-transform: y = H_0 H_1 ... H_{n-2} z   (apply last reflector first)
        for (int jj = 0; jj < kc; ++jj) {
          double* z = B.Z + jj * n;
          for (int k = n - 3; k >= 0; --k) {
            double tk = B.tau[k];
            if (tk == 0.0) continue;
            const int len = n - k - 1;
            double part = 0.0;
            DG_FOR(i, len) part += (i == 0 ? 1.0 : B.W[(k + 1 + i) * n + k]) * z[k + 1 + i];
            double dt = c.sum(part) * tk;
            DG_FOR(i, len) z[k + 1 + i] -= dt * (i == 0 ? 1.0 : B.W[(k + 1 + i) * n + k]);
            c.sync();
          }
        }
        // Hm += sum_j (floor - lam_j) y_j y_j'
        DG_FOR(t, n * n) {
          int i = t / n, j = t - i * n;
          double acc = 0.0;
          for (int jj = 0; jj < kc; ++jj) acc += (floor_val - B.lam[j0 + jj]) * B.Z[jj * n + i] * B.Z[jj * n + j];
          Hm[t] += acc;
        }
        c.sync();
      }
    }
  }
  if (reg > 0.0) { DG_FOR(i, n) Hm[i * n + i] += reg; }
  c.sync();
  return nneg;
}

// In-place lower Cholesky of the symmetric matrix Hm (row-major).  Returns false (uniformly) on a
// non-positive pivot.
DG_DEVN bool cholesky_lower(Cta& c, int n, double* Hm) {
  for (int k = 0; k < n; ++k) {
    double piv = Hm[k * n + k];
    if (!(piv > 0.0)) return false;
    double lkk = sqrt(piv), inv = 1.0 / lkk;
    c.sync();                                   // everyone has read the pivot
    for (int i = k + c.tid; i < n; i += c.nt) Hm[i * n + k] = i == k ? lkk : Hm[i * n + k] * inv;
    c.sync();
    for (int j = k + 1 + c.tid; j < n; j += c.nt) {
      double ljk = Hm[j * n + k];
      for (int i = j; i < n; ++i) Hm[i * n + j] -= Hm[i * n + k] * ljk;
    }
    c.sync();
  }
  return true;
}

// Jm = L^{-T}  (upper triangular, Jm Jm' = H^{-1}); thread per column of L^{-1}
DG_DEVN void tri_inverse_T(Cta& c, int n, const double* Lm, double* Jm) {
  DG_FOR(cc, n) {
    double* row = Jm + cc * n;
    for (int i = 0; i < cc; ++i) row[i] = 0.0;
    row[cc] = 1.0 / Lm[cc * n + cc];
    for (int i = cc + 1; i < n; ++i) {
      double acc = 0.0;
      for (int j = cc; j < i; ++j) acc += Lm[i * n + j] * row[j];
      row[i] = -acc / Lm[i * n + i];
    }
  }
  c.sync();
}
