"""OSQP restated (ADMM + polish) for the LITERAL mode of the oracle.  Oracle-only; test infrastructure.

The reference hands every QP sub-problem to OSQP through CasADi's ``conic`` plugin with ``polish=True`` and OSQP's
defaults for everything else (``DGSQP/solvers/DGSQP.py:186,246-249``).  OSQP itself is a third-party dependency that is
absent from ``/root/reference`` and not installable offline (``setup.py:14`` lists ``osqp`` unpinned; the solve actually
runs through the copy bundled in the unpinned ``casadi`` wheel), so this module restates its PUBLISHED algorithm
(Stellato, Banjac, Goulart, Bemporad, Boyd: "OSQP: an operator splitting solver for quadratic programs", Math. Prog.
Comp. 12, 2020; sections 3 (ADMM steps), 5.1 (Ruiz equilibration + cost scaling), 5.2 (adaptive rho), 3.4 (termination),
4 (polish)) at the upstream default settings

    rho = 0.1, sigma = 1e-6, alpha = 1.6, eps_abs = eps_rel = 1e-3, max_iter = 4000, scaling = 10,
    check_termination = 25, adaptive_rho = True (tolerance 5), delta = 1e-6, polish_refine_iter = 3.

One default cannot be restated: ``adaptive_rho_interval = 0`` makes OSQP pick the interval from WALL-CLOCK timing
(the first update happens once the iterations have taken 40 % of the setup time), which makes the reference itself
run-to-run irreproducible.  ``adaptive_rho_interval`` is therefore a parameter here (multiples of
``check_termination``; 25, 50 and 100 are what the rule typically lands on for problems of this size).

Purpose: ``scripts/literal_mode_study.py`` runs the oracle in literal mode (this solver, SciPy's plain ``lsqr``,
``mu_vio_thresh = 0``) against the product mode (exact QP, re-orthogonalised LSQR, ``mu_vio_thresh = 1e-10``) and reports
how many instances change ``(msg, num_iters)`` -- the only bound on the deviations D1-D3 available offline.

Problem form (CasADi's conic -> OSQP, ``DGSQP.py:246``):  min 1/2 x'Px + q'x  s.t.  A x <= u  (l = -inf).
"""
import numpy as np
import scipy.linalg as sla

OSQP_INFTY = 1e30
MIN_SCALING, MAX_SCALING = 1e-4, 1e4
RHO_MIN, RHO_MAX = 1e-6, 1e6


def _limit_scaling(v):
    v = np.where(v < MIN_SCALING, 1.0, v)
    return np.minimum(v, MAX_SCALING)


def _ruiz(P, q, A, iters=10):
    """Section 5.1: modified Ruiz equilibration of the KKT matrix [[P, A'], [A, 0]] plus the cost scaling c."""
    n, m = P.shape[0], A.shape[0]
    D, E, c = np.ones(n), np.ones(m), 1.0
    P, q, A = P.copy(), q.copy(), A.copy()
    for _ in range(iters):
        col_P = np.abs(P).max(axis=0)
        col_A = np.abs(A).max(axis=0) if m else np.zeros(n)
        dn = np.maximum(col_P, col_A)
        em = np.abs(A).max(axis=1) if m else np.zeros(0)
        dn = 1.0 / np.sqrt(_limit_scaling(dn))
        em = 1.0 / np.sqrt(_limit_scaling(em))
        P = dn[:, None] * P * dn[None, :]
        A = em[:, None] * A * dn[None, :]
        q = dn * q
        D *= dn
        E *= em
        # cost scaling
        cP = np.abs(P).max(axis=0).mean()
        cq = np.abs(q).max()
        cP = _limit_scaling(np.array([cP]))[0]
        cq = _limit_scaling(np.array([cq]))[0]
        ctmp = 1.0 / max(cP, cq)
        P *= ctmp
        q *= ctmp
        c *= ctmp
    return P, q, A, D, E, c


class OsqpResult:
    __slots__ = ("x", "y", "status", "iters", "polished", "rho_updates")


def solve_osqp(P, q, A, u, rho=0.1, sigma=1e-6, alpha=1.6, eps_abs=1e-3, eps_rel=1e-3, max_iter=4000, scaling=10,
               check_termination=25, adaptive_rho=True, adaptive_rho_interval=50, adaptive_rho_tolerance=5.0,
               polish=True, delta=1e-6, polish_refine_iter=3):
    """min 1/2 x'Px + q'x  s.t.  A x <= u.  Returns x, y (multipliers of A x <= u, >= 0 at a solution) as the ``x`` /
    ``lam_a`` CasADi hands back; cold start x = z = y = 0 like the reference (``DGSQP.py:240-241``: ``x0 = 0`` and no
    ``lam_a0``)."""
    n, m = P.shape[0], A.shape[0]
    Ps, qs, As, D, E, c = _ruiz(P, q, A, scaling) if scaling else (P.copy(), q.copy(), A.copy(), np.ones(n), np.ones(m), 1.0)
    us = np.where(u >= OSQP_INFTY, OSQP_INFTY, E * u)
    Einv, Dinv, cinv = 1.0 / E, 1.0 / D, 1.0 / c

    def factor(rho_):
        K = np.block([[Ps + sigma * np.eye(n), As.T], [As, -np.eye(m) / rho_]])
        return sla.lu_factor(K)          # quasi-definite; LU with pivoting is what a dense restatement can rely on

    lu = factor(rho)
    x, z, y = np.zeros(n), np.zeros(m), np.zeros(m)
    res = OsqpResult()
    res.status, res.polished, res.rho_updates = "max_iter", False, 0
    it = 0
    pri_res = dua_res = np.inf
    for it in range(1, max_iter + 1):
        rhs = np.concatenate([sigma * x - qs, z - y / rho])
        sol = sla.lu_solve(lu, rhs)
        xt, nu = sol[:n], sol[n:]
        zt = z + (nu - y) / rho
        x_new = alpha * xt + (1 - alpha) * x
        zr = alpha * zt + (1 - alpha) * z
        z_new = np.minimum(zr + y / rho, us)                   # projection onto (-inf, u]
        y = y + rho * (zr - z_new)
        x, z = x_new, z_new
        check = check_termination and it % check_termination == 0
        adapt = adaptive_rho and adaptive_rho_interval and it % adaptive_rho_interval == 0
        if check or adapt:
            Ax, Px, Aty = As @ x, Ps @ x, As.T @ y
            # unscaled residuals and tolerances (scaled_termination = False)
            pri_res = np.abs(Einv * (Ax - z)).max() if m else 0.0
            dua_res = cinv * np.abs(Dinv * (Px + qs + Aty)).max()
            if check:
                eps_pri = eps_abs + eps_rel * max(np.abs(Einv * Ax).max(), np.abs(Einv * z).max()) if m else eps_abs
                eps_dua = eps_abs + eps_rel * cinv * max(np.abs(Dinv * Px).max(), np.abs(Dinv * Aty).max(), np.abs(Dinv * qs).max())
                if pri_res < eps_pri and dua_res < eps_dua:
                    res.status = "solved"
                    break
            if adapt:
                # section 5.2, computed on the scaled quantities like osqp's compute_rho_estimate
                pr = np.abs(Ax - z).max() if m else 0.0
                dr = np.abs(Px + qs + Aty).max()
                pn = max(np.abs(Ax).max(), np.abs(z).max()) if m else 0.0
                dn = max(np.abs(Px).max(), np.abs(Aty).max(), np.abs(qs).max())
                pr_n = pr / (pn + 1e-10)
                dr_n = dr / (dn + 1e-10)
                rho_new = float(np.clip(rho * np.sqrt(pr_n / (dr_n + 1e-10)), RHO_MIN, RHO_MAX))
                if rho_new > rho * adaptive_rho_tolerance or rho_new < rho / adaptive_rho_tolerance:
                    rho = rho_new
                    lu = factor(rho)
                    res.rho_updates += 1
    res.iters = it
    if polish and res.status == "solved" and m:
        # section 4: guess the active set from the ADMM iterate, solve the reduced KKT system with the +-delta
        # regularisation and iterative refinement; keep the polished point only if it improves both residuals
        act = (us - z) < y
        na = int(act.sum())
        Aa = As[act]
        K = np.block([[Ps, Aa.T], [Aa, np.zeros((na, na))]])
        Kreg = K + np.diag(np.concatenate([delta * np.ones(n), -delta * np.ones(na)]))
        rhs = np.concatenate([-qs, us[act]])
        try:
            lur = sla.lu_factor(Kreg)
            t = sla.lu_solve(lur, rhs)
            for _ in range(polish_refine_iter):
                t = t + sla.lu_solve(lur, rhs - K @ t)
            xp, ya = t[:n], t[n:]
            yp = np.zeros(m)
            yp[act] = ya
            zp = As @ xp
            pol_pri = np.abs(Einv * np.maximum(zp - us, 0.0)).max()
            pol_dua = cinv * np.abs(Dinv * (Ps @ xp + qs + As.T @ yp)).max()
            if (pol_pri < pri_res and pol_dua < dua_res) or (pol_pri < pri_res and dua_res < 1e-10) or \
                    (pol_dua < dua_res and pri_res < 1e-10):
                x, y = xp, yp
                res.polished = True
        except Exception:
            pass
    res.x = D * x
    res.y = cinv * (E * y)
    return res
