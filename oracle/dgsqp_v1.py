"""DG-SQP v1 ``solve()``.  Oracle-only restatement of ``DGSQP/solvers/DGSQP.py``.

``solve``                :302-507        ``_solve_qp``            :232-266
``_get_mu``              :559-585        merit f_phi/f_dphi       :962-979
``_line_search_3``       :1057-1081      ``_watchdog_line_search_4`` :1174-1288
``_nearestPD``           :1290-1296

Deviations from the reference (documented in DESIGN.md):
* the QP is solved exactly (Goldfarb-Idnani) instead of by OSQP+polish; a QP the exact
  solver proves infeasible returns ``qp_fail`` (OSQP would hand NaNs to v1, which never
  checks ``success``);
* ``_get_mu`` compares the summed violation with ``thresh = 0`` (:560-569).  After a full QP step the
  active *linear* rows (input bounds, rate limits) evaluate to +-1 ulp, so with 0 the penalty weight flips
  between 0 and ~1e18 on rounding noise (in the reference: on OSQP's polish residual) and with it the
  whole iteration path.  ``mu_vio_thresh`` (default 1e-10) makes that choice deterministic; pass 0 for
  the literal reference rule;
* the dual initialisation runs LSQR with reorthogonalised Golub-Kahan vectors (``dual_init_method='reorth'``):
  SciPy's plain LSQR is only reproducible to ~1e-2 (oracle/lsqr_reorth.py); ``'scipy'`` gives the literal call;
* ``time_limit`` is not modelled (wall-clock, default None in every BASELINE config).
"""
import numpy as np
import scipy.sparse.linalg as spla

from .qp import solve_qp_gi, QPFailure
from .lsqr_reorth import lsqr_reorth


def nearest_pd(A, floor=1e-10):
    """_nearestPD (DGSQP.py:1290-1296)."""
    B = (A + A.T) / 2
    s, U = np.linalg.eigh(B)
    s[np.where(s < 0)[0]] = floor
    C = U @ np.diag(s) @ U.T
    return (C + C.T) / 2


class OracleDGSQP:
    def __init__(self, game, reg=1e-3, line_search_iters=50, nonmono_ls=True, sqp_iters=50,
                 p_tol=1e-3, d_tol=1e-3, beta=0.01, tau=0.5, merit_function="stat_l1",
                 conv_approx=True, mu_vio_thresh=1e-10, dual_init_method="reorth", qp_method="gi", osqp_kw=None,
                 l0_perturb=0.0):
        self.game = game
        self.reg, self.line_search_iters, self.nonmono_ls = reg, line_search_iters, nonmono_ls
        self.sqp_iters, self.p_tol, self.d_tol, self.beta, self.tau = sqp_iters, p_tol, d_tol, beta, tau
        self.merit_function, self.conv_approx = merit_function, conv_approx
        self.rel_tol_req = 3
        # `thresh` in _get_mu (DGSQP.py:560) is 0 in the reference; see module docstring
        self.mu_vio_thresh = mu_vio_thresh
        self.dual_init_method = dual_init_method        # 'reorth' (canonical) or 'scipy' (the literal call)
        # 'gi': exact KKT point (canonical, deviation D1); 'osqp': OSQP restated at its defaults with polish
        # (oracle/osqp_admm.py) -- the literal mode; like v1 it uses whatever iterate comes back (DGSQP.py:246-249)
        self.qp_method = qp_method
        self.osqp_kw = dict(osqp_kw or {})
        # sensitivity studies only (scripts/literal_mode_study.py): relative perturbation of the dual initialisation,
        # e.g. 2.2e-16 = one ulp; 0 in every parity run
        self.l0_perturb = l0_perturb
        self.qp_stats = []            # per-QP diagnostics (active-set size, negative eigenvalues, ...)
        self.trace = None

    # ------------------------------------------------------------------ pieces
    def _solve_qp(self, Q, q, G, g):
        if self.conv_approx:
            B = (Q + Q.T) / 2
            n_neg = int((np.linalg.eigvalsh(B) < 0).sum())
            H = nearest_pd(Q)
        else:
            n_neg = -1
            H = (Q + Q.T) / 2
        if self.reg > 0:
            H = H + self.reg * np.eye(H.shape[0])
        st = dict(n_neg=n_neg)
        if self.qp_method == "osqp":
            from .osqp_admm import solve_osqp
            r = solve_osqp(H, q, G, -g, **self.osqp_kw)
            st.update(osqp_status=r.status, iters=r.iters, polished=r.polished, rho_updates=r.rho_updates)
            self.qp_stats.append(st)
            if not (np.all(np.isfinite(r.x)) and np.all(np.isfinite(r.y))):
                return None, None
            return r.x, r.y
        try:
            du, l_hat = solve_qp_gi(H, q, G, g, stats=st)
        except QPFailure as e:
            st["fail"] = str(e)
            self.qp_stats.append(st)
            return None, None
        self.qp_stats.append(st)
        return du, l_hat

    def _phi(self, l, s, q, G, g, mu):
        stat = q + G.T @ l
        val = 0.5 * (stat @ stat + (l @ g) ** 2)
        if self.merit_function == "stat_l1":
            val += mu * np.sum(g - s)
        return val

    def _dstat_norm(self, du, l, dl, Q, q, G, g):
        d = q + G.T @ l
        lg = l @ g
        return d @ (Q @ du + G.T @ dl) + lg * (l @ (G @ du) + dl @ g)

    def _dphi(self, du, l, dl, s, Q, q, G, g, mu):
        val = self._dstat_norm(du, l, dl, Q, q, G, g)
        if self.merit_function == "stat_l1":
            val -= mu * np.sum(g - s)
        return val

    def _get_mu(self, du, l, dl, s, Q, q, G, g):
        if self.merit_function != "stat_l1":
            return 0.0
        vio = np.sum(g - s)
        if vio > self.mu_vio_thresh:
            return abs(self._dstat_norm(du, l, dl, Q, q, G, g)) / ((1 - 0.5) * vio)
        return 0.0

    def _line_search_3(self, u, du, l, dl, s, ds, Q, q, G, g, ev, mu):
        phi = self._phi(l, s, q, G, g, mu)
        dphi = self._dphi(du, l, dl, s, Q, q, G, g, mu)
        alpha = 1.0
        for _ in range(self.line_search_iters):
            u_t, l_t, s_t = u + alpha * du, l + alpha * dl, s + alpha * ds
            q_t, G_t, g_t, _ = ev(u_t, l_t, False)
            phi_t = self._phi(l_t, s_t, q_t, G_t, g_t, mu)
            self.n_ls_evals += 1
            if phi_t <= phi + self.beta * alpha * dphi:
                break
            alpha *= self.tau
        return u_t, l_t, phi_t

    def _watchdog_4(self, u_k, du_k, l_k, dl_k, s_k, ds_k, Q_k, q_k, G_k, g_k, ev, mu, merit_max=1e6):
        qp_solves = 0
        t_hat = 5
        phi_k = self._phi(l_k, s_k, q_k, G_k, g_k, mu)
        dphi_k = self._dphi(du_k, l_k, dl_k, s_k, Q_k, q_k, G_k, g_k, mu)
        u1, l1, s1 = u_k + du_k, l_k + dl_k, s_k + ds_k
        q1, G1, g1, _ = ev(u1, l1, False)
        phi1 = self._phi(l1, s1, q1, G1, g1, mu)
        if phi1 <= phi_k + self.beta * dphi_k:
            return u1, l1, qp_solves
        self.n_watchdog += 1
        fail = False
        u_t, l_t = u1, l1
        for _ in range(t_hat):
            Q_t, q_t, G_t, g_t, _ = ev(u_t, l_t, True)
            du_t, l_hat = self._solve_qp(Q_t, q_t, G_t, g_t)
            qp_solves += 1
            if du_t is None:
                fail = True
                break
            s_t = np.minimum(0, g_t)
            ds_t = g_t + G_t @ du_t - s_t
            u_n, l_n, s_n = u_t + du_t, l_hat, s_t + ds_t
            q_n, G_n, g_n, _ = ev(u_n, l_n, False)
            phi_n = self._phi(l_n, s_n, q_n, G_n, g_n, mu)
            if phi_n > merit_max:
                break
            if phi_n <= phi_k + self.beta * dphi_k:
                return u_n, l_n, qp_solves
            u_t, l_t = u_n, l_n
        # insist on merit decrease
        Q_t, q_t, G_t, g_t, _ = ev(u_t, l_t, True)
        du_t, l_hat = self._solve_qp(Q_t, q_t, G_t, g_t)
        qp_solves += 1
        if du_t is None:
            fail = True
        else:
            dl_t = l_hat - l_t
            s_t = np.minimum(0, g_t)
            ds_t = g_t + G_t @ du_t - s_t
            u_n, l_n, phi_n = self._line_search_3(u_t, du_t, l_t, dl_t, s_t, ds_t, Q_t, q_t, G_t, g_t, ev, mu)
        if not fail:
            if phi_n <= phi_k + self.beta * dphi_k:
                return u_n, l_n, qp_solves
            elif phi_n > phi_k:
                fail = True
            else:
                Q_n, q_n, G_n, g_n, _ = ev(u_n, l_n, True)
                du_n, l_hat = self._solve_qp(Q_n, q_n, G_n, g_n)
                if du_n is None:
                    u1, l1, _ = self._line_search_3(u_k, du_k, l_k, dl_k, s_k, ds_k, Q_k, q_k, G_k, g_k, ev, mu)
                    return u1, l1, qp_solves
                qp_solves += 1
                dl_n = l_hat - l_n
                s_n = np.minimum(0, g_n)
                ds_n = g_n + G_n @ du_n - s_n
                u2, l2, _ = self._line_search_3(u_n, du_n, l_n, dl_n, s_n, ds_n, Q_n, q_n, G_n, g_n, ev, mu)
                return u2, l2, qp_solves
        u1, l1, _ = self._line_search_3(u_k, du_k, l_k, dl_k, s_k, ds_k, Q_k, q_k, G_k, g_k, ev, mu)
        return u1, l1, qp_solves

    # -------------------------------------------------------------------- solve
    def dual_init(self, q, G):
        """l0 = max(0, -lsqr(G G^T, G q)) (DGSQP.py:323-324), SciPy's default tolerances.  'scipy' is the
        reference's literal call; 'reorth' is the same algorithm with reorthogonalised Golub-Kahan vectors,
        the reproducible form the CUDA path implements (see oracle/lsqr_reorth.py)."""
        if self.dual_init_method == "scipy":
            sol = spla.lsqr(G @ G.T, G @ q)
            x, self.lsqr_iters = sol[0], sol[2]
        else:
            mv = lambda v: G @ (G.T @ v)
            x, _, self.lsqr_iters = lsqr_reorth(mv, mv, G @ q, G.shape[0])
        return np.maximum(0, -x)

    def solve(self, x0, u_ws, record_trace=False, l_ws=None):
        game = self.game
        u = np.array(u_ws, dtype=np.float64).copy()
        up = np.zeros(game.n_u)                       # v1 zeroes u_prev every solve (:305)
        x0 = np.asarray(x0, dtype=np.float64)
        self.qp_stats, self.n_ls_evals, self.n_watchdog = [], 0, 0

        def ev(u_, l_, hessian):
            return game.evaluate(u_, l_, x0, up, hessian=hessian)

        if l_ws is None:
            q, G, _, _ = ev(u, None, False)
            l = self.dual_init(q, G)
            if self.l0_perturb:
                l = l * (1.0 + self.l0_perturb * np.random.default_rng(12345).uniform(-1.0, 1.0, l.shape))
        else:
            # commented-out alternative of the reference (DGSQP.py:313-319): caller-supplied duals
            l = np.array(l_ws, dtype=np.float64).copy()
        init = dict(u=u.copy(), l=l.copy())
        trace = [] if record_trace else None

        rel_tol_its, sqp_it, total_qp = 0, 0, 0
        converged = False
        while True:
            qp_solves = 0
            Q_i, q_i, G_i, g_i, _ = ev(u, l, True)
            d_i = q_i + G_i.T @ l
            u_im1, l_im1 = u.copy(), l.copy()
            p_feas = max(0, np.amax(g_i))
            comp = np.linalg.norm(g_i * l, ord=np.inf)
            stat = np.linalg.norm(d_i, ord=np.inf)
            cond = dict(p_feas=p_feas, comp=comp, stat=stat)
            if trace is not None:
                trace.append(dict(u=u.copy(), l=l.copy(), **cond))
            if stat > 1e5:
                msg = "diverged"
                break
            if p_feas < self.p_tol and comp < self.d_tol and stat < self.d_tol:
                converged, msg = True, "conv_abs_tol"
                break
            du, l_hat = self._solve_qp(Q_i, q_i, G_i, g_i)
            qp_solves += 1
            if du is None:
                total_qp += qp_solves
                msg = "qp_fail"
                break
            dl = l_hat - l
            s = np.minimum(0, g_i)
            ds = g_i + G_i @ du - s
            mu = self._get_mu(du, l, dl, s, Q_i, q_i, G_i, g_i)
            if self.nonmono_ls:
                u, l, n_qp = self._watchdog_4(u, du, l, dl, s, ds, Q_i, q_i, G_i, g_i, ev, mu)
                qp_solves += n_qp
            else:
                u, l, _ = self._line_search_3(u, du, l, dl, s, ds, Q_i, q_i, G_i, g_i, ev, mu)
            total_qp += qp_solves
            if np.linalg.norm(u - u_im1) < self.p_tol / 2 and np.linalg.norm(l - l_im1) < self.d_tol / 2:
                rel_tol_its += 1
                if rel_tol_its >= self.rel_tol_req and p_feas < self.p_tol:
                    converged, msg = True, "conv_rel_tol"
                    break
            else:
                rel_tol_its = 0
            sqp_it += 1
            if sqp_it >= self.sqp_iters:
                msg = "max_it"
                break
        x_bar = game.rollout(u, x0)
        J = game.costs(x_bar, u, up)
        self.trace = trace
        return dict(num_iters=sqp_it, status=converged, msg=msg, cost=J, cond=cond, init=init,
                    u=u, l=l, x=x_bar, qp_solves=total_qp)
