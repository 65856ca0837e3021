"""Racing dynamic game (chicane / curve / multi-agent).  Oracle-only restatement.

Game data follow ``scripts/DGSQP_ALGAMES_monte_carlo_chicane.py:49-127,222-330``,
``scripts/DGSQP_ALGAMES_monte_carlo_curve.py`` (same game, other literals) and
``scripts/DGSQP_monte_carlo_agents.py:47-231`` (M agents).  The vehicle model is
``CasadiKinematicBicycleCombined`` (``DGSQP/dynamics/dynamics_models.py:997-1079``)
under explicit Euler (``:90-91``); the joint model is block diagonal
(``:2482-2528``).  The condensed SQP quantities (x, A, B, Du_x, g, G, q, Q) restate
``DGSQP/solvers/DGSQP.py:587-979`` (``_build_solver``) and ``:509-533``
(``_evaluate``).
"""
import math
import numpy as np

from .jet import Jet

NQA, NUA = 6, 2          # per-agent state / input size: q=[x,y,v,e_psi,s,e_y], u=[a,delta]

# constraint row kinds
COLL, RATE, IN_UB, IN_LB, ST_UB, ST_LB = range(6)


class RacingGame:
    def __init__(self, track, M=2, N=25, dt=0.1, L_f=0.13, L_r=0.13, c_dr=0.1, c_da=0.0, c_s=0.1,
                 mass=2.366, input_weight=(1.0, 1.0), rate_weight=(1.0, 1.0), comp_weights=(10.0, 5.0),
                 u_ub=(2.1, 0.436), u_lb=(-2.1, -0.436), rate_ub=(10.0, math.pi), rate_lb=(-10.0, -math.pi),
                 half_width=1.0, obs_r=0.4):
        self.track, self.M, self.N, self.dt = track, M, N, dt
        self.L_f, self.L_r, self.c_dr, self.c_da, self.c_s, self.mass = L_f, L_r, c_dr, c_da, c_s, mass
        self.w_u = np.asarray(input_weight, float)
        self.w_du = np.asarray(rate_weight, float)
        self.c_prog, self.c_comp = comp_weights
        self.u_ub, self.u_lb = np.asarray(u_ub, float), np.asarray(u_lb, float)
        self.rate_ub, self.rate_lb = np.asarray(rate_ub, float), np.asarray(rate_lb, float)
        self.half_width = half_width
        self.obs_r = [obs_r] * M if np.isscalar(obs_r) else list(obs_r)
        self.n_q, self.n_u = NQA * M, NUA * M
        self.n = N * self.n_u
        self.pairs = [(i, j) for i in range(M) for j in range(i + 1, M)]
        # agent-major index of stage-major joint input (DGSQP.py:170)
        self.ua_idxs = [np.concatenate([np.arange(self.n_u * k + NUA * a, self.n_u * k + NUA * (a + 1))
                                        for k in range(N)]) for a in range(M)]
        self.perm = np.concatenate(self.ua_idxs)        # agent-major position -> stage-major index
        self._build_layout()

    # ------------------------------------------------------------------ layout
    def _build_layout(self):
        """Row order (DGSQP.py:730-821): stage major; within a stage
        [shared, agent0(user rate rows, in-ub, in-lb, (k>0) st-ub, st-lb), agent1, ...]."""
        M, N = self.M, self.N
        rows = []          # (stage, kind, a, b)   b: pair partner / component / rate row id
        self.n_c = []
        for k in range(N + 1):
            n0 = len(rows)
            if k >= 1:
                for (i, j) in self.pairs:
                    rows.append((k, COLL, i, j))
            for a in range(M):
                if k < N:
                    for r in range(4):
                        rows.append((k, RATE, a, r))
                    for c in range(NUA):
                        rows.append((k, IN_UB, a, c))
                    for c in range(NUA):
                        rows.append((k, IN_LB, a, c))
                if k > 0:
                    rows.append((k, ST_UB, a, 5))
                    rows.append((k, ST_LB, a, 5))
            self.n_c.append(len(rows) - n0)
        self.rows = rows
        self.m = len(rows)
        self.row_stage = np.array([r[0] for r in rows])
        self.row_kind = np.array([r[1] for r in rows])

    def uidx(self, a, k, c):
        """agent-major index of component c of u^a_k (DGSQP.py:590)."""
        return a * self.N * NUA + k * NUA + c

    # ---------------------------------------------------------------- dynamics
    def _fc_scalar(self, q, u):
        """Continuous dynamics of one agent, plain floats (dynamics_models.py:1046-1070)."""
        x, y, v, epsi, s, ey = q
        a, delta = u
        L = self.L_f + self.L_r
        beta = math.atan2(math.tan(delta) * self.L_r, L)
        psidot = v / self.L_r * math.sin(beta)
        absv = v if v > 0 else -v
        F = -self.c_da * v - self.c_dr * v * absv - self.c_s * psidot * psidot
        kap = float(self.track.curvature(s))
        psit = float(self.track.tangent(s)[0])
        den = 1.0 - ey * kap
        cb = math.cos(beta + epsi)
        return np.array([v * math.cos(beta + psit + epsi), v * math.sin(beta + psit + epsi),
                         a + F / self.mass, psidot - kap * v * cb / den, v * cb / den,
                         v * math.sin(beta + epsi)])

    def fc_joint(self, q, u):
        return np.concatenate([self._fc_scalar(q[NQA * a:NQA * (a + 1)], u[NUA * a:NUA * (a + 1)])
                               for a in range(self.M)])

    def rollout(self, u, x0):
        """evaluate_dynamics (DGSQP.py:598-601): x_{k+1} = x_k + dt*fc(x_k, u_k)."""
        N, M = self.N, self.M
        x = np.zeros((N + 1, self.n_q))
        x[0] = x0
        for k in range(N):
            uk = np.concatenate([u[self.uidx(a, k, 0):self.uidx(a, k, 0) + NUA] for a in range(M)])
            x[k + 1] = x[k] + self.dt * self.fc_joint(x[k], uk)
        return x

    def _fd_jets(self, xq, uu, order):
        """Jets of the discrete map for a batch of (state, input) pairs of single agents.
        xq: (K,6), uu: (K,2).  Returns list of 6 jets over variables [x,y,v,epsi,s,ey,a,delta]."""
        var = Jet.variables(np.hstack([xq, uu]), order=order)
        x, y, v, epsi, s, ey, a, delta = var
        L = self.L_f + self.L_r
        # atan2(tan(delta)*L_r, L) with L > 0 is atan(tan(delta)*L_r/L)
        beta = (delta.tan() * (self.L_r / L)).atan()
        psidot = v * (1.0 / self.L_r) * beta.sin()
        F = -self.c_da * v - self.c_dr * v * v.abs_ifelse() - self.c_s * psidot.square()
        kap = self.track.curvature(s.v)                # pw_const: zero derivative
        psit_v, psit_d = self.track.tangent(s.v)       # pw_lin: slope of the active segment
        # psi_t as a jet: value psit_v, gradient psit_d * ds, zero second derivative of the lookup
        psit = Jet(psit_v, psit_d[:, None] * s.g, None if s.h is None else psit_d[:, None, None] * s.h)
        den = 1.0 - ey * kap
        cb = (beta + epsi).cos()
        sb = (beta + epsi).sin()
        ang = beta + psit + epsi
        dq = [v * ang.cos(), v * ang.sin(), a + F * (1.0 / self.mass),
              psidot - (v * cb / den) * kap, v * cb / den, v * sb]
        return [var[i] + self.dt * dq[i] for i in range(NQA)]

    def linearize(self, x, u, order=2):
        """A_k, B_k (and second derivatives) at every stage for every agent.
        Returns Aa[N,M,6,6], Ba[N,M,6,2], H[N,M,6,8,8] (or None)."""
        N, M = self.N, self.M
        xq = x[:N].reshape(N * M, NQA)
        uu = np.stack([[u[self.uidx(a, k, 0):self.uidx(a, k, 0) + NUA] for a in range(M)]
                       for k in range(N)]).reshape(N * M, NUA)
        jets = self._fd_jets(xq, uu, order)
        Jg = np.stack([j.g for j in jets], axis=1).reshape(N, M, NQA, NQA + NUA)
        Aa, Ba = Jg[..., :NQA], Jg[..., NQA:]
        H = None
        if order >= 2:
            H = np.stack([j.h for j in jets], axis=1).reshape(N, M, NQA, NQA + NUA, NQA + NUA)
        return Aa, Ba, H

    def joint_AB(self, Aa, Ba, k):
        M = self.M
        A = np.zeros((self.n_q, self.n_q))
        B = np.zeros((self.n_q, self.n_u))
        for a in range(M):
            A[NQA * a:NQA * (a + 1), NQA * a:NQA * (a + 1)] = Aa[k, a]
            B[NQA * a:NQA * (a + 1), NUA * a:NUA * (a + 1)] = Ba[k, a]
        return A, B

    def sensitivities(self, Aa, Ba):
        """f_Du_x (DGSQP.py:642-650): S[k] = d x_k / d u, columns agent-major."""
        N = self.N
        S = np.zeros((N + 1, self.n_q, self.n))
        for k in range(N):
            A, B = self.joint_AB(Aa, Ba, k)
            S[k + 1] = A @ S[k]
            cols = self.perm_inv[self.n_u * k:self.n_u * (k + 1)]
            S[k + 1][:, cols] += B
        return S

    @property
    def perm_inv(self):
        if not hasattr(self, "_perm_inv"):
            inv = np.empty_like(self.perm)
            inv[self.perm] = np.arange(self.n)
            self._perm_inv = inv           # stage-major index -> agent-major position
        return self._perm_inv

    # ------------------------------------------------------------------ costs
    def costs(self, x, u, up):
        """f_J (DGSQP.py:889-893)."""
        N, M = self.N, self.M
        J = np.zeros(M)
        for a in range(M):
            ua = u[a * N * NUA:(a + 1) * N * NUA].reshape(N, NUA)
            um = np.vstack([up[NUA * a:NUA * (a + 1)], ua[:-1]])
            J[a] = 0.5 * np.sum(self.w_u * ua ** 2) + 0.5 * np.sum(self.w_du * (ua - um) ** 2)
            sN = x[N, NQA * a + 4]
            J[a] += -self.c_prog * sN
            for b in range(M):
                if b != a:
                    J[a] += self.c_comp * math.atan(x[N, NQA * b + 4] - sN)
        return J

    def _term_grad_hess(self, xN, a):
        """Gradient / Hessian wrt joint x_N of agent a's terminal cost
        -c0*s_a + sum_{b!=a} c1*atan(s_b - s_a)."""
        gx = np.zeros(self.n_q)
        Hx = np.zeros((self.n_q, self.n_q))
        ia = NQA * a + 4
        gx[ia] -= self.c_prog
        for b in range(self.M):
            if b == a:
                continue
            ib = NQA * b + 4
            d = xN[ib] - xN[ia]
            d1 = self.c_comp / (1.0 + d * d)
            d2 = -2.0 * self.c_comp * d / (1.0 + d * d) ** 2
            gx[ib] += d1
            gx[ia] -= d1
            Hx[ib, ib] += d2
            Hx[ia, ia] += d2
            Hx[ia, ib] -= d2
            Hx[ib, ia] -= d2
        return gx, Hx

    # ------------------------------------------------------------- constraints
    def constraints(self, x, u, up):
        """f_Cxu (DGSQP.py:804-821,911): g[m]."""
        N = self.N
        g = np.zeros(self.m)
        for r, (k, kind, a, b) in enumerate(self.rows):
            if kind == COLL:
                d = x[k, NQA * a:NQA * a + 2] - x[k, NQA * b:NQA * b + 2]
                g[r] = (self.obs_r[a] + self.obs_r[b]) ** 2 - d @ d
            elif kind == RATE:
                c = b // 2
                uk = u[self.uidx(a, k, c)]
                um = up[NUA * a + c] if k == 0 else u[self.uidx(a, k - 1, c)]
                du = uk - um
                g[r] = du - self.dt * self.rate_ub[c] if b % 2 == 0 else self.dt * self.rate_lb[c] - du
            elif kind == IN_UB:
                g[r] = u[self.uidx(a, k, b)] - self.u_ub[b]
            elif kind == IN_LB:
                g[r] = self.u_lb[b] - u[self.uidx(a, k, b)]
            elif kind == ST_UB:
                g[r] = x[k, NQA * a + b] - self.half_width
            elif kind == ST_LB:
                g[r] = -self.half_width - x[k, NQA * a + b]
        return g

    def constraint_jacobian(self, x, S):
        """f_Du_C (DGSQP.py:824-826,918): G = dC/du + dC/dx * Du_x, dense (m, n)."""
        G = np.zeros((self.m, self.n))
        for r, (k, kind, a, b) in enumerate(self.rows):
            if kind == COLL:
                d = x[k, NQA * a:NQA * a + 2] - x[k, NQA * b:NQA * b + 2]
                G[r] = -2.0 * d @ (S[k][NQA * a:NQA * a + 2] - S[k][NQA * b:NQA * b + 2])
            elif kind == RATE:
                c = b // 2
                sg = 1.0 if b % 2 == 0 else -1.0
                G[r, self.uidx(a, k, c)] = sg
                if k > 0:
                    G[r, self.uidx(a, k - 1, c)] = -sg
            elif kind == IN_UB:
                G[r, self.uidx(a, k, b)] = 1.0
            elif kind == IN_LB:
                G[r, self.uidx(a, k, b)] = -1.0
            elif kind == ST_UB:
                G[r] = S[k][NQA * a + b]
            elif kind == ST_LB:
                G[r] = -S[k][NQA * a + b]
        return G

    def _constraint_lx_lxx(self, x, l):
        """Stage-wise gradient / Hessian wrt x_k of sum_j l_j C_j (state-dependent rows only)."""
        N = self.N
        lx = np.zeros((N + 1, self.n_q))
        lxx = np.zeros((N + 1, self.n_q, self.n_q))
        for r, (k, kind, a, b) in enumerate(self.rows):
            if kind == COLL:
                ia, ib = NQA * a, NQA * b
                d = x[k, ia:ia + 2] - x[k, ib:ib + 2]
                lx[k, ia:ia + 2] += -2.0 * l[r] * d
                lx[k, ib:ib + 2] += 2.0 * l[r] * d
                for c in range(2):
                    lxx[k, ia + c, ia + c] += -2.0 * l[r]
                    lxx[k, ib + c, ib + c] += -2.0 * l[r]
                    lxx[k, ia + c, ib + c] += 2.0 * l[r]
                    lxx[k, ib + c, ia + c] += 2.0 * l[r]
            elif kind == ST_UB:
                lx[k, NQA * a + b] += l[r]
            elif kind == ST_LB:
                lx[k, NQA * a + b] -= l[r]
        return lx, lxx

    # ------------------------------------------------------------ gradient q
    def cost_gradient(self, x, u, up, S):
        """f_q (DGSQP.py:673-676,898-899): q = [grad_{u^a} J^a]_a."""
        N, M = self.N, self.M
        q = np.zeros(self.n)
        for a in range(M):
            sl = slice(a * N * NUA, (a + 1) * N * NUA)
            ua = u[sl].reshape(N, NUA)
            um = np.vstack([up[NUA * a:NUA * (a + 1)], ua[:-1]])
            du = ua - um
            ga = self.w_u * ua + self.w_du * du
            ga[:-1] -= self.w_du * du[1:]
            gx, _ = self._term_grad_hess(x[N], a)
            q[sl] = ga.ravel() + gx @ S[N][:, sl]
        return q

    def sum_cost_gradient(self, x, u, up, S):
        """grad_u sum_a J^a (the `dobj` of the v2 merit 'sum_obj_l1', DGSQP_v2.py:1150-1151): agent a's inputs enter
        its own stage costs directly and every agent's terminal cost through the states."""
        N, M = self.N, self.M
        qs = self.cost_gradient(x, u, up, S)
        for a in range(M):
            sl = slice(a * N * NUA, (a + 1) * N * NUA)
            for f in range(M):
                if f != a:
                    gx, _ = self._term_grad_hess(x[N], f)
                    qs[sl] += gx @ S[N][:, sl]
        return qs

    # ------------------------------------------------------------- Hessian Q
    def _dp_hessian(self, Aa, Ba, H, lx, lxx, luu, luu2):
        """Backward dynamic-programming Hessian of  Phi(u) = sum_k l_k(x_k,u_k,u_{k-1}) + l_N(x_N)
        wrt the stage-major joint input sequence (DGSQP.py:679-727 for costs, :829-877 for
        constraints).  lx[N+1,nq], lxx[N+1,nq,nq]; luu[N,nu,nu] = d2(l_k+l_{k+1})/du_k^2;
        luu2[N,nu,nu] = d2 l_{k+1}/du_{k+1} du_k.  Stage costs here never couple x and u."""
        N, M, nq, nu = self.N, self.M, self.n_q, self.n_u
        p = lx[N].copy()
        V = lxx[N].copy()
        W = np.zeros((0, nq))
        Duu = np.zeros((0, 0))
        for k in range(N - 1, -1, -1):
            A, B = self.joint_AB(Aa, Ba, k)
            # sum_i p_i * Hess(fd_i): block diagonal over agents (E: xx, F: uu, G: ux)
            E = np.zeros((nq, nq))
            F = np.zeros((nu, nu))
            Gm = np.zeros((nu, nq))
            for a in range(M):
                Hc = np.tensordot(p[NQA * a:NQA * (a + 1)], H[k, a], axes=(0, 0))   # (8,8)
                E[NQA * a:NQA * (a + 1), NQA * a:NQA * (a + 1)] = Hc[:NQA, :NQA]
                F[NUA * a:NUA * (a + 1), NUA * a:NUA * (a + 1)] = Hc[NQA:, NQA:]
                Gm[NUA * a:NUA * (a + 1), NQA * a:NQA * (a + 1)] = Hc[NQA:, :NQA]
            A1 = luu[k] + B.T @ V @ B + F
            if W.shape[0] == 0:
                Duu_k = A1
            else:
                B1 = W @ B
                B1[:nu] += luu2[k]
                Duu_k = np.block([[A1, B1.T], [B1, Duu]])
            A2 = B.T @ V @ A + Gm
            W = A2 if W.shape[0] == 0 else np.vstack([A2, W @ A])
            V_new = lxx[k] + A.T @ V @ A + E
            p = lx[k] + p @ A
            V = V_new
            Duu = Duu_k
        # stage-major -> agent-major (DGSQP.py:725-726)
        return Duu[np.ix_(self.perm, self.perm)]

    def hessian(self, x, u, l, Aa, Ba, H):
        """f_Q (DGSQP.py:920-934): row block a = grad_{u^a} grad_u (J^a + l^T C)."""
        N, M, nq, nu = self.N, self.M, self.n_q, self.n_u
        Q = np.zeros((self.n, self.n))
        zero_lx = np.zeros((N + 1, nq))
        for a in range(M):
            lx = zero_lx.copy()
            lxx = np.zeros((N + 1, nq, nq))
            lx[N], lxx[N] = self._term_grad_hess(x[N], a)
            luu = np.zeros((N, nu, nu))
            luu2 = np.zeros((N, nu, nu))
            for k in range(N):
                for c in range(NUA):
                    i = NUA * a + c
                    luu[k, i, i] = self.w_u[c] + self.w_du[c] + (self.w_du[c] if k < N - 1 else 0.0)
                    if k < N - 1:
                        luu2[k, i, i] = -self.w_du[c]
            Da = self._dp_hessian(Aa, Ba, H, lx, lxx, luu, luu2)
            sl = slice(a * N * NUA, (a + 1) * N * NUA)
            Q[sl] = Da[sl]
        lx, lxx = self._constraint_lx_lxx(x, l)
        Q += self._dp_hessian(Aa, Ba, H, lx, lxx, np.zeros((N, nu, nu)), np.zeros((N, nu, nu)))
        return Q

    # ---------------------------------------------------------------- evaluate
    def evaluate(self, u, l, x0, up, hessian=True):
        """_evaluate (DGSQP.py:509-533).  Returns (Q, q, G, g, x) or (q, G, g, x)."""
        x = self.rollout(u, x0)
        Aa, Ba, H = self.linearize(x, u, order=2 if hessian else 1)
        S = self.sensitivities(Aa, Ba)
        g = self.constraints(x, u, up)
        G = self.constraint_jacobian(x, S)
        q = self.cost_gradient(x, u, up, S)
        # kept for the v2 merit 'sum_obj_l1' (sum of the costs and its gradient at the evaluated point)
        self.last_obj, self.last_qs = float(np.sum(self.costs(x, u, up))), self.sum_cost_gradient(x, u, up, S)
        if hessian:
            Q = self.hessian(x, u, l, Aa, Ba, H)
            return Q, q, G, g, x
        return q, G, g, x
