"""DG-SQP v2 ``solve()``.  Oracle-only restatement of ``DGSQP/solvers/DGSQP_v2.py``.

``solve``            :322-652     ``_solve_qp``        :253-285     ``_get_mu``        :683-707
``load_checkpoint``  :709-727     ``line_search``      :729-760     merit f_phi/f_dphi :1143-1167
``_nearest_pd``      :1269-1275   constants            :86 (rel_tol_req = 10), :205-213

v2 keeps v1's SQP approximation (same ``_evaluate``) and changes the step policy: a decaying
regularisation, relaxed "d-steps" accepted without a merit test while the step stays inside a shrinking radius
``delta``, "m-steps" that test the merit function against the maximum of a short memory and fall back to a line
search from the last checkpoint iterate, and ``max_it`` counted in m-steps.

Merit ('stat_l1', :1143-1161): ``stat = [grad_{u^a} L^a]_a = q + G'l``; ``phi = 1/2 |stat|^2 + mu * sum(s)`` with the
slack ``s = max(0, g)`` handed in by the caller; ``dphi = stat' (Q du + G' dl) - mu * sum(s)`` (the Jacobian of ``stat``
w.r.t. ``u`` is exactly the row-stacked game Hessian ``Q`` and w.r.t. ``l`` it is ``G'``).

Deviations (same as the v1 oracle): exact QP instead of OSQP (D1), ``mu_vio_thresh`` (D2),
re-orthogonalised LSQR dual initialisation (D3), no wall-clock ``time_limit``.

Merit 'sum_obj_l1' (:1149-1151,1161-1164): ``phi = sum_a J^a(u) + mu * sum(s)``, ``dphi = grad_u(sum_a J^a)' du -
mu * sum(s)``; the games keep both quantities of the last evaluated point (``last_obj``, ``last_qs``).
"""
from collections import deque

import numpy as np

from .dgsqp_v1 import OracleDGSQP, nearest_pd
from .qp import solve_qp_gi, QPFailure


class OracleDGSQPV2(OracleDGSQP):
    def __init__(self, game, reg=1e2, reg_decay=0.95, line_search_iters=50, nms=True, nms_frequency=5,
                 nms_memory_size=3, sqp_iters=500, p_tol=1e-4, d_tol=1e-4, beta=0.25, tau=0.5,
                 merit_function="stat_l1", merit_parameter=None, merit_decrease=0.01,
                 merit_decrease_condition="armijo", delta_decay=0.95, mu_vio_thresh=1e-10, dual_init_method="reorth"):
        if merit_function not in ("stat_l1", "sum_obj_l1"):
            raise ValueError(f"Merit function option {merit_function} not recognized")
        if merit_decrease_condition not in ("armijo", "max"):
            raise ValueError(f"Merit decrease condition {merit_decrease_condition} not recognized")
        super().__init__(game, reg=reg, line_search_iters=line_search_iters, sqp_iters=sqp_iters, p_tol=p_tol, d_tol=d_tol,
                         beta=beta, tau=tau, merit_function="stat_l1", mu_vio_thresh=mu_vio_thresh,
                         dual_init_method=dual_init_method)
        self.merit_obj = merit_function == "sum_obj_l1"
        self.reg_init, self.reg_decay = reg, reg_decay
        self.nms, self.nms_mstep_frequency, self.nms_memory_size = nms, nms_frequency, nms_memory_size
        self.merit_parameter, self.sigma, self.gamma = merit_parameter, merit_decrease, delta_decay
        self.merit_decrease_condition = merit_decrease_condition
        self.rel_tol_req = 10                            # :86
        self.nms_initial_step_size_factor = 20           # :212
        self.nms_initial_reference_factor = 1            # :213

    # ------------------------------------------------------------------ pieces
    def _solve_qp2(self, Q, q, G, g, l):
        """:253-285 -- eigenvalue floor 1e-9 (:1273), current regularisation, returns (du, dl)."""
        H = nearest_pd(Q, floor=1e-9)
        if self.reg > 0:
            H = H + self.reg * np.eye(H.shape[0])
        try:
            du, l_hat = solve_qp_gi(H, q, G, g)
        except QPFailure:
            return None, None
        return du, l_hat - l

    def _phi2(self, l, s, q, G, mu):
        """f_phi at the point the game was evaluated at last (every call site follows that evaluation)."""
        if self.merit_obj:
            return self.game.last_obj + mu * np.sum(s)
        stat = q + G.T @ l
        return 0.5 * (stat @ stat) + mu * np.sum(s)

    def _dstat2(self, du, l, dl, Q, q, G):
        """f_dphi_c: directional derivative of the merit's smooth part along (du, dl)."""
        if self.merit_obj:
            return float(self.game.last_qs @ du)
        return (q + G.T @ l) @ (Q @ du + G.T @ dl)

    def _get_mu2(self, du, l, dl, s, Q, q, G):
        """:683-707 -- all four cases reduce to |d_phi_c| / ((1 - rho) sum(s)) with violation, else 0."""
        d_phi_c = self._dstat2(du, l, dl, Q, q, G)
        vio = np.sum(s)
        return abs(d_phi_c) / ((1 - 0.5) * vio) if vio > self.mu_vio_thresh else 0.0

    def _line_search2(self, u, du, l, dl, s, mu, ev, memory):
        """:729-760"""
        if self.merit_decrease_condition == "max":
            ref = lambda a: (1 - self.sigma * a) * max(memory)
        else:
            Q0, q0, G0, g0, _ = ev(u, l, True)
            phi0 = self._phi2(l, s, q0, G0, mu)
            dphi0 = self._dstat2(du, l, dl, Q0, q0, G0) - mu * np.sum(np.maximum(0, g0))
            ref = lambda a: phi0 + self.sigma * a * dphi0
        a = 1.0
        for _ in range(self.line_search_iters):
            u_t, l_t = u + a * du, l + a * dl
            q_t, G_t, g_t, _ = ev(u_t, l_t, False)
            s_t = np.maximum(0, g_t)
            self.n_ls_evals += 1
            if self._phi2(l_t, s_t, q_t, G_t, mu) <= ref(a):
                break
            a *= self.tau
        return u_t, l_t, a, self._phi2(l_t, s_t, q_t, G_t, 1.0)

    # -------------------------------------------------------------------- solve
    def solve(self, x0, u_ws, u_prev=None, record_trace=False, l_ws=None):
        game = self.game
        u = np.array(u_ws, dtype=np.float64).copy()
        up = np.zeros(game.n_u) if u_prev is None else np.asarray(u_prev, dtype=np.float64)   # v2 keeps u_prev (:328)
        x0 = np.asarray(x0, dtype=np.float64)
        self.n_ls_evals = 0

        def ev(u_, l_, hessian):
            return game.evaluate(u_, l_, x0, up, hessian=hessian)

        q, G, g, _ = ev(u, None, False)
        l = self.dual_init(q, G) if l_ws is None else np.array(l_ws, dtype=np.float64).copy()
        init = dict(u=u.copy(), l=l.copy())
        u_im1, l_im1 = u.copy(), l.copy()
        memory = deque([self.nms_initial_reference_factor * self._phi2(l, np.maximum(0, g), q, G, 1.0)],
                       self.nms_memory_size)
        self.reg = self.reg_init
        ck_counter, ck_index, ck_delta, ck_reg = 0, 0, 0.0, self.reg
        delta = 0.0
        records = []              # per appended iteration: (u, du, l, dl, s, ds, mu) -- IterationData fields load_checkpoint reads
        trace = [] if record_trace else None

        converged, rel_tol_its, sqp_it, m_step_it, total_qp, n_dstep = False, 0, 0, 0, 0, 0
        finished, msg = False, None
        while True:
            rec_u, rec_l = u.copy(), l.copy()
            Q_i, q_i, G_i, g_i, _ = ev(u, l, True)
            d_i = q_i + G_i.T @ l
            p_feas = max(0, np.amax(g_i))
            comp = np.linalg.norm(g_i * l, ord=np.inf)
            stat = np.linalg.norm(d_i, ord=np.inf)
            cond = dict(p_feas=p_feas, comp=comp, stat=stat)
            if trace is not None:
                trace.append(dict(u=u.copy(), l=l.copy(), reg=self.reg, delta=delta, **cond))
            if stat > 1e10:                                   # :394 (later tests override the message, as in the reference)
                converged, finished, msg = False, True, "diverged"
            if p_feas < self.p_tol and comp < self.d_tol and stat < self.d_tol:
                converged, finished, msg = True, True, "conv_abs_tol"
            if m_step_it >= self.sqp_iters:                    # :407
                converged, finished, msg = False, True, "max_it"
            if finished:
                break

            du, dl = self._solve_qp2(Q_i, q_i, G_i, g_i, l)
            total_qp += 1
            rec = None
            if du is None:
                if not self.nms or len(records) == 0:          # :439-444, :459-464
                    msg = "qp_fail"
                    break
                d_step, m_step = False, True
                rec = records[min(ck_index, len(records) - 1)]  # :449
                u, du, l, dl, s, ds, mu = rec
            else:
                if sqp_it == 0:
                    delta = self.nms_initial_step_size_factor * np.linalg.norm(np.concatenate((du, dl)))
                    ck_delta = delta
                if self.nms:
                    d_step, m_step = False, False
                    if ck_counter >= self.nms_mstep_frequency:
                        m_step = True
                    elif np.linalg.norm(np.concatenate((du, dl))) < delta:
                        d_step = True
                    else:
                        m_step = True
                else:
                    d_step, m_step = False, False
                s = np.maximum(0, g_i)
                ds = np.maximum(0, g_i + G_i @ du) - s
                mu = self._get_mu2(du, l, dl, s, Q_i, q_i, G_i) if self.merit_parameter is None else self.merit_parameter
                rec = (rec_u, du.copy(), rec_l, dl.copy(), s.copy(), ds.copy(), mu)

            if d_step:
                u = u + du
                l = l + dl
                delta = self.gamma * delta
                ck_counter += 1
                n_dstep += 1
            if m_step:
                m_step_it += 1
                u_f, l_f = u + du, l + dl
                q_f, G_f, g_f, _ = ev(u_f, l_f, False)
                phi = self._phi2(l_f, np.maximum(0, g_f), q_f, G_f, 1.0)
                if phi <= (1 - self.sigma * 1.0) * max(memory):
                    u, l = u_f, l_f
                else:
                    if ck_index <= len(records) - 1:           # :533
                        u, du, l, dl, s, ds, mu = records[ck_index]
                        rec = records[ck_index]                 # load_checkpoint overwrites _data with the checkpoint's fields
                        delta = ck_delta
                        self.reg = ck_reg
                    u, l, _, phi = self._line_search2(u, du, l, dl, s, mu, ev, memory)
                if np.linalg.norm(u - u_im1) < self.p_tol and np.linalg.norm(l - l_im1) < self.d_tol:
                    rel_tol_its += 1
                    if rel_tol_its >= self.rel_tol_req and p_feas < self.p_tol:
                        converged, finished, msg = True, True, "conv_rel_tol"
                else:
                    rel_tol_its = 0
                u_im1, l_im1 = u.copy(), l.copy()
                self.reg = self.reg * self.reg_decay
                memory.append(phi)
                ck_counter, ck_delta, ck_reg, ck_index = 0, delta, self.reg, sqp_it + 1
            if (not d_step) and (not m_step):
                u, l, _, phi = self._line_search2(u, du, l, dl, s, mu, ev, memory)
                if np.linalg.norm(u - u_im1) < self.p_tol and np.linalg.norm(l - l_im1) < self.d_tol:
                    rel_tol_its += 1
                    if rel_tol_its >= self.rel_tol_req and p_feas < self.p_tol:
                        converged, finished, msg = True, True, "conv_rel_tol"
                else:
                    rel_tol_its = 0
                u_im1, l_im1 = u.copy(), l.copy()
                self.reg = self.reg * self.reg_decay
                memory.append(phi)
            records.append(rec)
            sqp_it += 1
            # a 'conv_rel_tol' raised above is only acted upon at the top of the next iteration (:415), after one more
            # evaluation whose tests can still override the message -- exactly as in the reference
        x_bar = game.rollout(u, x0)
        J = game.costs(x_bar, u, up)
        self.trace = trace
        return dict(num_iters=sqp_it, status=converged, msg=msg, cost=J, cond=cond, init=init, u=u, l=l, x=x_bar,
                    qp_solves=total_qp, m_steps=m_step_it, d_steps=n_dstep)
