// CTA-cooperative dense FP64 kernels for one n x n problem (n ~ 100-200), row-major storage.
//
// nearest_pd() replaces DGSQP._nearestPD (DGSQP/solvers/DGSQP.py:1290-1296: B=(A+A')/2; eigh;
// negative eigenvalues -> floor; U diag(s) U'; re-symmetrise).  Only the negative part of the
// spectrum changes, so   nearestPD(A) = B + sum_{s_i<0} (floor - s_i) u_i u_i'   and only the
// negative eigenpairs are needed: Householder tridiagonalisation, Sturm-count multisection for the
// negative eigenvalues, inverse iteration on the tridiagonal matrix, back-transformation.  When the
// Sturm count at 0 is zero the projection is the identity and the whole eigen-solve is skipped.
//
// Performance notes (B200): these routines are latency bound (n ~ 100 wide, ~n dependent steps), so
// they are written to (1) keep every global access coalesced across the CTA (thread = column),
// (2) expose instruction-level parallelism in the per-thread loops (explicit 4-way unrolling on
// __restrict__ pointers: read-modify-write loops otherwise serialise on possible aliasing), and
// (3) spend as few CTA barriers as possible (fused reductions, blocked Cholesky with the panel in
// shared memory, warp-private back-transformation).
#pragma once
#include "cta.cuh"

#define DG_EIG_CHUNK 16     // eigenvectors computed concurrently by inverse iteration
#define DG_CHOL_NB 8        // Cholesky panel width

struct LinBuf {
  double* W;      // n*n   tridiagonalisation workspace (holds the reflectors afterwards)
  double* dg;     // n     tridiagonal diagonal
  double* od;     // n     off-diagonal (od[k] couples k, k+1)
  double* od2;    // n     od^2
  double* tau;    // n
  double* lam;    // n     negative eigenvalues (ascending)
  double* Z;      // DG_EIG_CHUNK*n  eigenvectors of the chunk
  double* itw;    // DG_EIG_CHUNK*5*n  inverse-iteration factor storage
  // shared memory scratch
  double* pv;     // n
  double* wv;     // n
  double* sp;     // n*DG_CHOL_NB  Cholesky panel
};

// Householder reduction of the symmetric matrix W (full storage, both triangles kept consistent)
// to tridiagonal form; reflector k is stored in W[k+2.., k] (v[0] = 1 implicit) with tau[k].
// Three barriers per step: [trailing update + next norm], [matvec + dot], [w ready].
DG_DEVN void sym_tridiag(Cta& c, int n, const LinBuf& B) {
  double* DG_RESTRICT W = B.W;
  double* DG_RESTRICT pv = B.pv;
  double* DG_RESTRICT wv = B.wv;
  // norm of the first column below the sub-diagonal
  double part = 0.0;
  for (int i = c.tid + 2; i < n; i += c.nt) { double xv = W[i * n]; part += xv * xv; }
  double xn2 = c.sum(part);
  for (int k = 0; k + 1 < n; ++k) {
    const int len = n - k - 1;            // x = W[k+1.., k]
    const int off = k + 1;
    const double alpha = W[off * n + k];
    double tauk = 0.0, beta = alpha, scale = 0.0;
    if (xn2 > 0.0) {
      beta = -copysign(sqrt(alpha * alpha + xn2), alpha);
      tauk = (beta - alpha) / beta;
      scale = 1.0 / (alpha - beta);
    }
    if (c.tid == 0) { B.dg[k] = W[k * n + k]; B.od[k] = beta; B.od2[k] = beta * beta; B.tau[k] = tauk; }
    if (tauk == 0.0) {
      // nothing to annihilate: next column norm straight from memory
      part = 0.0;
      for (int i = c.tid + 2; i < len; i += c.nt) { double xv = W[(off + i) * n + off]; part += xv * xv; }
      xn2 = c.sum(part);
      continue;
    }
    // p = tau * A22 v with v_j = (j == 0 ? 1 : W[off+j][k] * scale); thread i owns column i (coalesced)
    double pdot = 0.0;
    DG_FOR(i, len) {
      const double* DG_RESTRICT col = W + off * n + off + i;
      const double* DG_RESTRICT vc = W + off * n + k;
      double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
      int j = 1;
      for (; j + 4 <= len; j += 4) {
        a0 += col[(j + 0) * n] * vc[(j + 0) * n];
        a1 += col[(j + 1) * n] * vc[(j + 1) * n];
        a2 += col[(j + 2) * n] * vc[(j + 2) * n];
        a3 += col[(j + 3) * n] * vc[(j + 3) * n];
      }
      for (; j < len; ++j) a1 += col[j * n] * vc[j * n];
      // only the j >= 1 terms carry the scale factor
      double acc = (col[0] + scale * ((a0 + a1) + (a2 + a3))) * tauk;            // j = 0 term has v_0 = 1
      double vi = i == 0 ? 1.0 : vc[i * n] * scale;
      pv[i] = vi;
      wv[i] = acc;
      pdot += acc * vi;
    }
    const double hk = 0.5 * tauk * c.sum(pdot);
    DG_FOR(i, len) {
      wv[i] -= hk * pv[i];
      if (i > 0) W[(off + i) * n + k] = pv[i];                    // keep the reflector
    }
    c.sync();
    // A22 -= v w' + w v'   (thread per column); accumulate the next column norm on the fly
    part = 0.0;
    DG_FOR(j, len) {
      const double vj = pv[j], wj = wv[j];
      double* DG_RESTRICT col = W + off * n + off + j;
      int i = 0;
      for (; i + 4 <= len; i += 4) {
        double c0 = col[(i + 0) * n], c1 = col[(i + 1) * n], c2 = col[(i + 2) * n], c3 = col[(i + 3) * n];
        c0 -= pv[i + 0] * wj + wv[i + 0] * vj;
        c1 -= pv[i + 1] * wj + wv[i + 1] * vj;
        c2 -= pv[i + 2] * wj + wv[i + 2] * vj;
        c3 -= pv[i + 3] * wj + wv[i + 3] * vj;
        col[(i + 0) * n] = c0; col[(i + 1) * n] = c1; col[(i + 2) * n] = c2; col[(i + 3) * n] = c3;
      }
      for (; i < len; ++i) col[i * n] -= pv[i] * wj + wv[i] * vj;
      // next step's x is column 0 of the updated block below its sub-diagonal == row 0, columns >= 2
      if (j >= 2) { double xv = col[0]; part += xv * xv; }
    }
    xn2 = c.sum(part);
  }
  if (c.tid == 0) { B.dg[n - 1] = W[(n - 1) * n + (n - 1)]; B.od[n - 1] = 0.0; B.od2[n - 1] = 0.0; }
  c.sync();
}

// number of eigenvalues of tridiag(dg, od) that are < x   (od2 = od^2)
DG_DEV int sturm_count(int n, const double* DG_RESTRICT dg, const double* DG_RESTRICT od2, double x, double pivmin) {
  int cnt = 0;
  double q = dg[0] - x;
  if (fabs(q) < pivmin) q = -pivmin;
  cnt += q < 0.0;
  for (int i = 1; i < n; ++i) {
    q = dg[i] - x - od2[i - 1] / q;
    if (fabs(q) < pivmin) q = -pivmin;
    cnt += q < 0.0;
  }
  return cnt;
}

// Solve (T - lam I) y = b in place (y overwrites b) by Gaussian elimination with partial pivoting.
// fw: 5*n scratch (diag, sup1, sup2, mult, swapped)
DG_DEV void tridiag_shift_solve(int n, const double* dg, const double* od, double lam, double tiny,
                                double* fw, double* y, bool refactor) {
  double* a = fw; double* b1 = fw + n; double* b2 = fw + 2 * n; double* ml = fw + 3 * n; double* sw = fw + 4 * n;
  if (refactor) {
    // work row "cur" = (p, q, r) starting at column i
    double p = dg[0] - lam, q = n > 1 ? od[0] : 0.0, r = 0.0;
    for (int i = 0; i + 1 < n; ++i) {
      double sub = od[i], nd = dg[i + 1] - lam, ns = i + 2 < n ? od[i + 1] : 0.0;
      if (fabs(sub) > fabs(p)) {
        double m = p / sub;                       // pivot row is the next row (sub, nd, ns)
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
  for (int i = 0; i + 1 < n; ++i) {
    if (sw[i] != 0.0) { double t = y[i]; y[i] = y[i + 1]; y[i + 1] = t - ml[i] * y[i]; }
    else y[i + 1] -= ml[i] * y[i];
  }
  for (int i = n - 1; i >= 0; --i) {
    double t = y[i];
    if (i + 1 < n) t -= b1[i] * y[i + 1];
    if (i + 2 < n) t -= b2[i] * y[i + 2];
    double piv = a[i];
    if (fabs(piv) < tiny) piv = piv < 0.0 ? -tiny : tiny;
    y[i] = t / piv;
  }
}

// Negative eigenvalues of the tridiagonal matrix into B.lam[0..nneg): Sturm-count multisection.  All
// eigenvalues are refined together: each round spends the CTA's nt probes evenly over the brackets.
DG_DEV void negative_eigenvalues(Cta& c, int n, const LinBuf& B, int nneg, double tnorm, double pivmin,
                                 double* lo, double* hi) {
  // lo/hi: shared or global scratch of nneg doubles each (B.pv / B.wv are free here)
  DG_FOR(j, nneg) { lo[j] = -tnorm * 1.0000001 - pivmin; hi[j] = 0.0; }
  c.sync();
  int per = c.nt / nneg;                       // probes per eigenvalue per round
  if (per < 1) per = 1;
  const int groups = c.nt / per;               // eigenvalues refined concurrently
  int rounds = (int)ceil(56.0 * 0.6931471805599453 / log((double)per + 1.0));
  if (rounds < 1) rounds = 1;
  int* cnts = (int*)(B.itw);                   // nt ints of scratch
  for (int j0 = 0; j0 < nneg; j0 += groups) {
    const int g = c.tid / per, t = c.tid - g * per, j = j0 + g;
    const bool act = g < groups && j < nneg;
    for (int rd = 0; rd < rounds; ++rd) {
      double a = 0.0, b = 0.0, h = 0.0;
      if (act) {
        a = lo[j]; b = hi[j];
        h = (b - a) / (double)(per + 1);
        cnts[c.tid] = sturm_count(n, B.dg, B.od2, a + h * (double)(t + 1), pivmin);
      }
      c.sync();
      if (act && t == 0) {
        int first = per;                        // first probe with count >= j+1
        for (int s = 0; s < per; ++s) if (cnts[c.tid + s] >= j + 1) { first = s; break; }
        lo[j] = first > 0 ? a + h * (double)first : a;
        hi[j] = first < per ? a + h * (double)(first + 1) : b;
      }
      c.sync();
    }
  }
  DG_FOR(j, nneg) B.lam[j] = 0.5 * (lo[j] + hi[j]);
  c.sync();
}

// Hm <- nearestPD(Qraw) + reg*I.  Returns the number of negative eigenvalues (uniform across threads).
DG_DEVN int nearest_pd(Cta& c, int n, const double* DG_RESTRICT Qraw, double* DG_RESTRICT Hm, const LinBuf& B,
                       double floor_val, double reg, bool conv_approx) {
  // symmetric part into W and Hm (thread = column: Qraw[i][j] coalesced, Qraw[j][i] strided but L1 resident)
  c.lap(PH_OTHER);
  DG_FOR(t, n * n) {
    int i = t / n, j = t - i * n;
    double sv = 0.5 * (Qraw[i * n + j] + Qraw[j * n + i]);
    B.W[t] = sv; Hm[t] = sv;
  }
  c.sync();
  int nneg = 0;
  if (conv_approx) {
    sym_tridiag(c, n, B);
    c.lap(PH_PD_TRIDIAG);
    double tn = 0.0;
    DG_FOR(i, n) {
      double r = fabs(B.dg[i]) + fabs(B.od[i]) + (i > 0 ? fabs(B.od[i - 1]) : 0.0);
      tn = fmax(tn, r);
    }
    const double tnorm = c.max(tn);
    const double pivmin = 2.2250738585072014e-308 * fmax(1.0, tnorm * tnorm);
    nneg = sturm_count(n, B.dg, B.od2, 0.0, pivmin);      // every thread computes the same count
    if (nneg > 0) {
      negative_eigenvalues(c, n, B, nneg, tnorm, pivmin, B.pv, B.wv);
      // --- eigenvectors in chunks: inverse iteration (thread per vector), Gram-Schmidt inside clusters,
      //     warp-private back-transformation, rank-one corrections of Hm
      const double tiny = fmax(tnorm, 1.0) * 1.1e-16;
      for (int j0 = 0; j0 < nneg; j0 += DG_EIG_CHUNK) {
        const int kc = nneg - j0 < DG_EIG_CHUNK ? nneg - j0 : DG_EIG_CHUNK;
        bool cluster = false;
        for (int jj = 1; jj < kc; ++jj)
          if (fabs(B.lam[j0 + jj] - B.lam[j0 + jj - 1]) <= 1e-3 * tnorm) cluster = true;
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
          if (cluster) {
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
        // back-transform y = H_0 H_1 ... H_{n-2} z (last reflector first): one warp per eigenvector,
        // no CTA barrier inside
        for (int jj = c.warp; jj < kc; jj += c.nwarps) {
          double* DG_RESTRICT z = B.Z + jj * n;
          for (int k = n - 3; k >= 0; --k) {
            const double tk = B.tau[k];
            if (tk == 0.0) continue;
            const int len = n - k - 1;
            const double* DG_RESTRICT vcol = B.W + (k + 1) * n + k;
            double p = 0.0;
            for (int i = c.lane; i < len; i += c.wsz) p += (i == 0 ? 1.0 : vcol[i * n]) * z[k + 1 + i];
            p = c.warp_sum(p) * tk;
            for (int i = c.lane; i < len; i += c.wsz) z[k + 1 + i] -= p * (i == 0 ? 1.0 : vcol[i * n]);
            c.syncwarp();
          }
        }
        c.sync();
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
  c.lap(PH_PD_EIG);
  return nneg;
}

// In-place lower Cholesky of the symmetric matrix Hm (row-major), blocked right-looking with the
// panel (n x NB) staged in shared memory.  Returns false (uniformly) on a non-positive pivot.
DG_DEVN bool cholesky_lower(Cta& c, int n, double* DG_RESTRICT Hm, double* DG_RESTRICT sp) {
  const int NB = DG_CHOL_NB;
  for (int k0 = 0; k0 < n; k0 += NB) {
    const int nb = n - k0 < NB ? n - k0 : NB;
    const int rows = n - k0;
    // stage panel rows k0.. into shared memory: sp[(i-k0)*NB + t] = Hm[i][k0+t]
    for (int e = c.tid; e < rows * NB; e += c.nt) {
      int r = e / NB, t = e - r * NB;
      sp[e] = t < nb ? Hm[(k0 + r) * n + k0 + t] : 0.0;
    }
    c.sync();
    for (int kk = 0; kk < nb; ++kk) {
      // pivot (every thread computes it from the finished row kk of the panel)
      double piv = sp[kk * NB + kk];
      for (int t = 0; t < kk; ++t) piv -= sp[kk * NB + t] * sp[kk * NB + t];
      if (!(piv > 0.0)) return false;
      const double inv = 1.0 / sqrt(piv);
      c.sync();                                   // all pivots read before row kk's entry is overwritten
      for (int r = kk + c.tid; r < rows; r += c.nt) {
        double v = sp[r * NB + kk];
        if (r == kk) v = piv * inv;               // sqrt(piv)
        else {
          for (int t = 0; t < kk; ++t) v -= sp[r * NB + t] * sp[kk * NB + t];
          v *= inv;
        }
        sp[r * NB + kk] = v;
      }
      c.sync();
    }
    // write the finished panel back (upper part of the diagonal block is left untouched: never read)
    for (int e = c.tid; e < rows * NB; e += c.nt) {
      int r = e / NB, t = e - r * NB;
      if (t < nb && t <= r) Hm[(k0 + r) * n + k0 + t] = sp[e];
    }
    // trailing update of the lower triangle: thread = column j, rows i >= j; panel rows from shared memory
    const int j0 = k0 + nb;
    for (int j = j0 + c.tid; j < n; j += c.nt) {
      double lj[DG_CHOL_NB];
#pragma unroll
      for (int t = 0; t < DG_CHOL_NB; ++t) lj[t] = sp[(j - k0) * NB + t];
      double* DG_RESTRICT col = Hm + j;
      int i = j;
      for (; i + 2 <= n; i += 2) {
        const double* DG_RESTRICT r0 = sp + (i - k0) * NB;
        const double* DG_RESTRICT r1 = r0 + NB;
        double c0 = col[i * n], c1 = col[(i + 1) * n];
#pragma unroll
        for (int t = 0; t < DG_CHOL_NB; ++t) { c0 -= r0[t] * lj[t]; c1 -= r1[t] * lj[t]; }
        col[i * n] = c0; col[(i + 1) * n] = c1;
      }
      for (; i < n; ++i) {
        const double* DG_RESTRICT r0 = sp + (i - k0) * NB;
        double c0 = col[i * n];
#pragma unroll
        for (int t = 0; t < DG_CHOL_NB; ++t) c0 -= r0[t] * lj[t];
        col[i * n] = c0;
      }
    }
    c.sync();
  }
  return true;
}

// Y = L^{-1} (lower triangular, row-major: Y[i][c]); thread c owns column c, so every access is coalesced.
//   Y[c][c] = 1/L[c][c];  Y[i][c] = -(sum_{j=c}^{i-1} L[i][j] Y[j][c]) / L[i][i]
// The GI solver uses J = L^{-T} = Y' through the accessor J(i,j) = Y[j*n+i].
DG_DEVN void tri_inverse(Cta& c, int n, const double* DG_RESTRICT Lm, double* DG_RESTRICT Y) {
  DG_FOR(cc, n) {
    for (int i = 0; i < cc; ++i) Y[i * n + cc] = 0.0;
    Y[cc * n + cc] = 1.0 / Lm[cc * n + cc];
    for (int i = cc + 1; i < n; ++i) {
      const double* DG_RESTRICT Li = Lm + i * n;
      double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
      int j = cc;
      for (; j + 4 <= i; j += 4) {
        a0 += Li[j] * Y[j * n + cc];
        a1 += Li[j + 1] * Y[(j + 1) * n + cc];
        a2 += Li[j + 2] * Y[(j + 2) * n + cc];
        a3 += Li[j + 3] * Y[(j + 3) * n + cc];
      }
      for (; j < i; ++j) a0 += Li[j] * Y[j * n + cc];
      Y[i * n + cc] = -((a0 + a1) + (a2 + a3)) / Li[i];
    }
  }
  c.sync();
}
