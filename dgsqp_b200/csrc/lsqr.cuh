// Dual initialisation  l0 = max(0, -lsqr(G G', G q))   (DGSQP/solvers/DGSQP.py:320-326).
//
// The reference calls SciPy's LSQR (Paige & Saunders) with its defaults atol = btol = 1e-6,
// conlim = 1e8, iter_lim = 2*m, damp = 0, x0 = None on the singular m x m matrix G G'.  The
// early-stopped Krylov iterate is not a closed form, so the recurrences and the ORDER of the
// stopping tests are restated here after scipy/sparse/linalg/_isolve/lsqr.py:329-560.
//
// Plain LSQR loses orthogonality of its Golub-Kahan vectors and then amplifies rounding: SciPy run on
// the dense and on the sparse form of the same G G' already returns l0 that differ by up to 5e-3 and
// stops up to one iteration apart (measured).  To make the dual initialisation a
// reproducible function of the instance, every new u_k / v_k is re-orthogonalised against all earlier
// ones (classical Gram-Schmidt, two passes; basis capped at DG_LSQR_BASIS vectors).  In exact arithmetic
// this is the same algorithm; in FP64 the iterates then track the exact-arithmetic ones to ~1e-12.
#pragma once
#include "game.cuh"

#define DG_LSQR_BASIS 64

struct LsqrBuf {
  double* Ub;  // DG_LSQR_BASIS*m  stored u_k
  double* Vb;  // DG_LSQR_BASIS*m  stored v_k
  double* cf;  // DG_LSQR_BASIS    projection coefficients
  double* u;   // m
  double* v;   // m
  double* w;   // m
  double* x;   // m
  double* tn;  // n  scratch (G' v)
  double* tm;  // m  scratch (A v)
};

DG_DEV void sym_ortho(double a, double b, double& cs, double& sn, double& r) {
  if (b == 0.0) { cs = a > 0.0 ? 1.0 : (a < 0.0 ? -1.0 : 0.0); sn = 0.0; r = fabs(a); }
  else if (a == 0.0) { cs = 0.0; sn = b > 0.0 ? 1.0 : -1.0; r = fabs(b); }
  else if (fabs(b) > fabs(a)) {
    double tau = a / b;
    sn = (b > 0.0 ? 1.0 : -1.0) / sqrt(1.0 + tau * tau);
    cs = sn * tau; r = b / sn;
  } else {
    double tau = b / a;
    cs = (a > 0.0 ? 1.0 : -1.0) / sqrt(1.0 + tau * tau);
    sn = cs * tau; r = a / cs;
  }
}

// out = G (G' in)
template <bool SM>
DG_DEV void lsqr_A(Cta& c, const Dims& D, const EvalBuf& E, const LsqrBuf& L, const double* in, double* out) {
  game_GT_times<SM>(c, D, E, in, L.tn);
  game_G_times<SM>(c, D, E, L.tn, out);
}

template <bool SM>
DG_DEV double vec_norm(Cta& c, int len, const double* v) {
  double p = 0.0;
  DG_FOR(i, len) p += v[i] * v[i];
  return sqrt(c.sum(p));
}

// vec <- (I - B B') vec, twice; B = first nb rows of basis (orthonormal).  Warp per basis vector for the
// dot products, thread per element for the update.
template <bool SM>
DG_DEV void lsqr_reorth(Cta& c, int m, const double* basis, int nb, double* vec, double* cf) {
  for (int pass = 0; pass < 2; ++pass) {
    c.sync();
    for (int j = c.warp(); j < nb; j += c.nwarps()) {
      const double* b = basis + (size_t)j * m;
      double p = 0.0;
      for (int i = c.lane(); i < m; i += c.wsz) p += b[i] * vec[i];
      p = c.warp_sum(p);
      if (c.lane() == 0) cf[j] = p;
    }
    c.sync();
    DG_FOR(i, m) {
      double acc = 0.0;
      for (int j = 0; j < nb; ++j) acc += cf[j] * basis[(size_t)j * m + i];
      vec[i] -= acc;
    }
  }
  c.sync();
}

// l_out[m] = max(0, -x_lsqr).  Returns the iteration count.
template <bool SM>
DG_DEVN int lsqr_dual_init(Cta& c, const Dims& D_, const EvalBuf& E_, const LsqrBuf& L_, const double* qv, double* l_out) {
  // local copies: the tables live in shared memory and would otherwise be re-read after every store
  const EvalBuf E = E_; DG_SH_EVAL(E); const LsqrBuf L = L_; const Dims D = D_;
  const int m = D.m;
  const double atol = 1e-6, btol = 1e-6, conlim = 1e8, eps = 2.220446049250313e-16;
  const int iter_lim = 2 * m;
  const double ctol = 1.0 / conlim;
  int itn = 0, istop = 0;
  double anorm = 0, acond = 0, ddnorm = 0, res2 = 0, xnorm = 0, xxnorm = 0, z = 0, cs2 = -1, sn2 = 0;
  // b = G q -> u
  game_G_times<SM>(c, D, E, qv, L.u);
  DG_FOR(i, m) L.x[i] = 0.0;
  double bnorm = vec_norm<SM>(c, m, L.u);
  double beta = bnorm, alfa = 0.0;
  int nU = 0, nV = 0;
  if (beta > 0.0) {
    c.sync();
    DG_FOR(i, m) { L.u[i] *= 1.0 / beta; L.Ub[i] = L.u[i]; }
    nU = 1;
    c.sync();
    lsqr_A<SM>(c, D, E, L, L.u, L.v);
    alfa = vec_norm<SM>(c, m, L.v);
  } else {
    DG_FOR(i, m) L.v[i] = 0.0;
  }
  c.sync();
  if (alfa > 0.0) { DG_FOR(i, m) { L.v[i] *= 1.0 / alfa; L.Vb[i] = L.v[i]; } nV = 1; }
  c.sync();
  DG_FOR(i, m) L.w[i] = L.v[i];
  double rhobar = alfa, phibar = beta;
  double arnorm = alfa * beta;
  if (arnorm != 0.0) {
    while (itn < iter_lim) {
      ++itn;
      // u = A v - alfa u
      lsqr_A<SM>(c, D, E, L, L.v, L.tm);
      DG_FOR(i, m) L.u[i] = L.tm[i] - alfa * L.u[i];
      lsqr_reorth<SM>(c, m, L.Ub, nU, L.u, L.cf);
      beta = vec_norm<SM>(c, m, L.u);
      if (beta > 0.0) {
        c.sync();
        const bool keepU = nU < DG_LSQR_BASIS;
        DG_FOR(i, m) { L.u[i] *= 1.0 / beta; if (keepU) L.Ub[(size_t)nU * m + i] = L.u[i]; }
        if (keepU) ++nU;
        c.sync();
        anorm = sqrt(anorm * anorm + alfa * alfa + beta * beta);
        lsqr_A<SM>(c, D, E, L, L.u, L.tm);
        DG_FOR(i, m) L.v[i] = L.tm[i] - beta * L.v[i];
        lsqr_reorth<SM>(c, m, L.Vb, nV, L.v, L.cf);
        alfa = vec_norm<SM>(c, m, L.v);
        c.sync();
        if (alfa > 0.0) {
          const bool keepV = nV < DG_LSQR_BASIS;
          DG_FOR(i, m) { L.v[i] *= 1.0 / alfa; if (keepV) L.Vb[(size_t)nV * m + i] = L.v[i]; }
          if (keepV) ++nV;
        }
        c.sync();
      }
      double rhobar1 = rhobar, psi = 0.0;
      double cs, sn, rho;
      sym_ortho(rhobar1, beta, cs, sn, rho);
      double theta = sn * alfa;
      rhobar = -cs * alfa;
      double phi = cs * phibar;
      phibar = sn * phibar;
      double tau = sn * phi;
      double t1 = phi / rho, t2 = -theta / rho;
      double dpart = 0.0;
      DG_FOR(i, m) {
        double wi = L.w[i];
        double dk = (1.0 / rho) * wi;
        dpart += dk * dk;
        L.x[i] += t1 * wi;
        L.w[i] = L.v[i] + t2 * wi;
      }
      double dkn = sqrt(c.sum(dpart));
      ddnorm += dkn * dkn;
      double delta = sn2 * rho, gambar = -cs2 * rho, rhs = phi - delta * z, zbar = rhs / gambar;
      xnorm = sqrt(xxnorm + zbar * zbar);
      double gamma = sqrt(gambar * gambar + theta * theta);
      cs2 = gambar / gamma; sn2 = theta / gamma; z = rhs / gamma;
      xxnorm += z * z;
      acond = anorm * sqrt(ddnorm);
      double res1 = phibar * phibar;
      res2 += psi * psi;
      double rnorm = sqrt(res1 + res2);
      arnorm = alfa * fabs(tau);
      double test1 = rnorm / bnorm;
      double test2 = arnorm / (anorm * rnorm + eps);
      double test3 = 1.0 / (acond + eps);
      double tt1 = test1 / (1.0 + anorm * xnorm / bnorm);
      double rtol = btol + atol * anorm * xnorm / bnorm;
#ifdef DG_LSQR_DEBUG
      printf("%3d rnorm %.6e test1 %.6e rtol %.6e test2 %.3e anorm %.6e acond %.4e xnorm %.6e alfa %.6e beta %.6e\n", itn, rnorm, test1, rtol, test2, anorm, acond, xnorm, alfa, beta);
#endif
      if (itn >= iter_lim) istop = 7;
      if (1.0 + test3 <= 1.0) istop = 6;
      if (1.0 + test2 <= 1.0) istop = 5;
      if (1.0 + tt1 <= 1.0) istop = 4;
      if (test3 <= ctol) istop = 3;
      if (test2 <= atol) istop = 2;
      if (test1 <= rtol) istop = 1;
      if (istop != 0) break;
    }
  }
  c.sync();
  DG_FOR(i, m) { double v = -L.x[i]; l_out[i] = v > 0.0 ? v : 0.0; }
  c.sync();
  return itn;
}
