"""LSQR with full reorthogonalisation of the Golub-Kahan vectors.  Oracle-only.

Same recurrences, scalars and stopping tests as ``scipy.sparse.linalg.lsqr`` (Paige & Saunders; the call
the reference makes at ``DGSQP/solvers/DGSQP.py:324`` with SciPy's defaults), but every new u_k / v_k is
re-orthogonalised (modified Gram-Schmidt, twice) against all previous ones.  In exact arithmetic this
changes nothing.  In floating point plain LSQR loses orthogonality once Ritz values converge and then
amplifies rounding errors: two FP64 implementations of the *same* recurrences (SciPy on a dense vs. a
sparse matrix, this oracle vs. the CUDA kernel) stop up to 2 iterations apart and return l0 that differ by
1e-4..1e-2.  With reorthogonalisation the iterates follow the exact-arithmetic ones to ~1e-12, so the dual
initialisation becomes a well-defined function of the instance that independent implementations agree on.
"""
import math
import numpy as np


def _sym_ortho(a, b):
    if b == 0:
        return np.sign(a), 0.0, abs(a)
    if a == 0:
        return 0.0, np.sign(b), abs(b)
    if abs(b) > abs(a):
        tau = a / b
        s = np.sign(b) / math.sqrt(1 + tau * tau)
        return s * tau, s, b / s
    tau = b / a
    c = np.sign(a) / math.sqrt(1 + tau * tau)
    return c, c * tau, a / c


def _reorth(vec, basis):
    for _ in range(2):
        for b in basis:
            vec = vec - (b @ vec) * b
    return vec


def lsqr_reorth(matvec, rmatvec, b, n, atol=1e-6, btol=1e-6, conlim=1e8, iter_lim=None, max_basis=64):
    m = b.shape[0]
    if iter_lim is None:
        iter_lim = 2 * n
    eps = np.finfo(np.float64).eps
    ctol = 1.0 / conlim if conlim > 0 else 0.0
    itn = istop = 0
    anorm = acond = ddnorm = res2 = xnorm = xxnorm = z = 0.0
    cs2, sn2 = -1.0, 0.0
    u = b.copy()
    bnorm = np.linalg.norm(b)
    x = np.zeros(n)
    beta = bnorm
    Us, Vs = [], []
    if beta > 0:
        u = u / beta
        Us.append(u)
        v = rmatvec(u)
        alfa = np.linalg.norm(v)
    else:
        v = x.copy()
        alfa = 0.0
    if alfa > 0:
        v = v / alfa
        Vs.append(v)
    w = v.copy()
    rhobar, phibar = alfa, beta
    arnorm = alfa * beta
    if arnorm == 0:
        return x, 0, 0
    while itn < iter_lim:
        itn += 1
        u = matvec(v) - alfa * u
        u = _reorth(u, Us)
        beta = np.linalg.norm(u)
        if beta > 0:
            u = u / beta
            if len(Us) < max_basis:
                Us.append(u)
            anorm = math.sqrt(anorm ** 2 + alfa ** 2 + beta ** 2)
            v = rmatvec(u) - beta * v
            v = _reorth(v, Vs)
            alfa = np.linalg.norm(v)
            if alfa > 0:
                v = v / alfa
                if len(Vs) < max_basis:
                    Vs.append(v)
        rhobar1 = rhobar
        cs, sn, rho = _sym_ortho(rhobar1, beta)
        theta = sn * alfa
        rhobar = -cs * alfa
        phi = cs * phibar
        phibar = sn * phibar
        tau = sn * phi
        t1, t2 = phi / rho, -theta / rho
        dk = (1 / rho) * w
        x = x + t1 * w
        w = v + t2 * w
        ddnorm += np.linalg.norm(dk) ** 2
        delta = sn2 * rho
        gambar = -cs2 * rho
        rhs = phi - delta * z
        zbar = rhs / gambar
        xnorm = math.sqrt(xxnorm + zbar ** 2)
        gamma = math.sqrt(gambar ** 2 + theta ** 2)
        cs2, sn2 = gambar / gamma, theta / gamma
        z = rhs / gamma
        xxnorm += z ** 2
        acond = anorm * math.sqrt(ddnorm)
        rnorm = math.sqrt(phibar ** 2 + res2)
        arnorm = alfa * abs(tau)
        test1 = rnorm / bnorm
        test2 = arnorm / (anorm * rnorm + eps)
        test3 = 1 / (acond + eps)
        tt1 = test1 / (1 + anorm * xnorm / bnorm)
        rtol = btol + atol * anorm * xnorm / bnorm
        if itn >= iter_lim:
            istop = 7
        if 1 + test3 <= 1:
            istop = 6
        if 1 + test2 <= 1:
            istop = 5
        if 1 + tt1 <= 1:
            istop = 4
        if test3 <= ctol:
            istop = 3
        if test2 <= atol:
            istop = 2
        if test1 <= rtol:
            istop = 1
        if istop:
            break
    return x, istop, itn
