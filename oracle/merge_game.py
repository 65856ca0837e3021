"""Highway-merge dynamic game (three unicycles, one on a ramp).  Oracle-only restatement.

Game data follow ``scripts/DGSQP_merge_monte_carlo.py``: lane geometry ``:40-74``, vehicle models
``:87-127`` (``CasadiKinematicUnicycle``, ``DGSQP/dynamics/dynamics_models.py:306-345``, RK3 with one
sub-step ``:202-212``, joint block-diagonal model ``:2482-2528``), bounds ``:130-161``, collision
radii ``:164-169``, costs ``:253-304``, constraints ``:306-398``, solver parameters ``:178-192``.
The condensed SQP quantities (x, A, B, Du_x, g, G, q, Q) restate ``DGSQP/solvers/DGSQP.py:587-979``.

This file is written independently of ``oracle/racing_game.py`` (dense joint matrices, a generic stage-cost
DP) so that the two restatements of the reference's Hessian recursion check each other.

The script builds its models from a ``DynamicsConfig`` without ``mass``; the value used is the
``DynamicsConfig`` default 2.366 (``DGSQP/dynamics/model_types.py:99``).
"""
import math
import numpy as np

from .jet import Jet

NQA, NUA = 4, 2          # per-agent state / input: q = [x, y, v, psi], u = [F_x, w_z]

# constraint row kinds
COLL, LANE, IN_UB, IN_LB, ST_UB, ST_LB = range(6)


def merge_lanes(lw=0.3, mw=0.3, mp=1.5, th=math.pi / 12):
    """Lane half-planes per agent (merge.py:40-74,316-318).  A lane row is
    ``n(x)'(p - (pt - r*n(x))) <= 0`` with ``n(x) = n_a`` for ``x < brk`` and ``n_b`` otherwise
    (CasADi ``pw_const``).  Returns [[(brk, n_a, n_b, pt), (..)] for each of the three cars]."""
    ns = np.array([0.0, 1.0])
    nm = np.array([-math.sin(th), math.cos(th)])
    x1 = np.array([0.0, lw])
    x3 = np.array([0.0, 0.0])
    x6 = np.array([mp + lw / math.tan(th), lw])
    x7 = np.array([mp + mw / math.sin(th), 0.0])
    inf = float("inf")
    straight = [(inf, ns, ns, x1), (inf, -ns, -ns, x3)]
    ramp = [(float(x6[0]), nm, ns, x6), (float(x7[0]), -nm, -ns, x7)]
    return [straight, straight, ramp]


class MergeGame:
    def __init__(self, M=3, N=20, dt=0.1, mass=2.366, input_weight=(0.1, 0.1), state_weight=(1.0, 10.0, 1.0, 1.0),
                 term_scale=10.0, goals=((4.0, 0.15, 0.3, 0.0), (4.5, 0.15, 0.3, 0.0), (4.25, 0.15, 0.3, 0.0)),
                 u_ub=(2.0, 4.5), u_lb=(-2.0, -4.5), v_ub=2.0, v_lb=-2.0, obs_r=(0.1, 0.1, 0.1), lanes=None,
                 lane_r=0.1):
        self.M, self.N, self.dt, self.mass = M, N, dt, mass
        self.w_u = np.asarray(input_weight, float)
        self.w_q = np.asarray(state_weight, float)
        self.term_scale = term_scale
        self.goals = np.asarray(goals, float)
        self.u_ub, self.u_lb = np.asarray(u_ub, float), np.asarray(u_lb, float)
        self.v_ub, self.v_lb = v_ub, v_lb
        self.obs_r = list(obs_r)
        self.lanes = merge_lanes() if lanes is None else lanes
        self.lane_r = lane_r
        assert len(self.goals) == M and len(self.obs_r) == M and len(self.lanes) == M
        self.n_q, self.n_u = NQA * M, NUA * M
        self.n = N * self.n_u
        self.pairs = [(i, j) for i in range(M) for j in range(i + 1, M)]
        # agent-major <-> stage-major input ordering (DGSQP.py:170,590)
        self.perm = np.array([self.n_u * k + NUA * a + c for a in range(M) for k in range(N) for c in range(NUA)])
        self.perm_inv = np.empty_like(self.perm)
        self.perm_inv[self.perm] = np.arange(self.n)
        self._build_layout()

    # ------------------------------------------------------------------ layout
    def _build_layout(self):
        """Row order (DGSQP.py:730-821): stage major; within a stage [shared (k>=1), then per agent:
        user rows (2 lane rows, every k), in-ub, in-lb (k<N), st-ub, st-lb (k>0; only v is bounded)]."""
        rows, self.n_c = [], []
        for k in range(self.N + 1):
            n0 = len(rows)
            if k >= 1:
                rows += [(k, COLL, i, j) for (i, j) in self.pairs]
            for a in range(self.M):
                rows += [(k, LANE, a, 0), (k, LANE, a, 1)]
                if k < self.N:
                    rows += [(k, IN_UB, a, c) for c in range(NUA)]
                    rows += [(k, IN_LB, a, c) for c in range(NUA)]
                if k > 0:
                    rows += [(k, ST_UB, a, 2), (k, ST_LB, a, 2)]
            self.n_c.append(len(rows) - n0)
        self.rows, self.m = rows, len(rows)

    def uidx(self, a, k, c):
        return a * self.N * NUA + k * NUA + c

    def _u_stage(self, u, k):
        return np.concatenate([u[self.uidx(a, k, 0):self.uidx(a, k, 0) + NUA] for a in range(self.M)])

    # ---------------------------------------------------------------- dynamics
    def _fc(self, q, u):
        """Unicycle (dynamics_models.py:325-333) on arrays or jets: q, u are sequences of components."""
        x, y, v, psi = q
        F, w = u
        if isinstance(psi, Jet):
            return [v * psi.cos(), v * psi.sin(), F * (1.0 / self.mass), w]
        return [v * math.cos(psi), v * math.sin(psi), F / self.mass, w]

    def _rk3(self, q, u):
        """rk3 with M = 1 sub-step, h = dt (dynamics_models.py:202-212)."""
        h = self.dt
        a1 = [h * f for f in self._fc(q, u)]
        a2 = [h * f for f in self._fc([q[i] + a1[i] * 0.5 for i in range(NQA)], u)]
        a3 = [h * f for f in self._fc([q[i] - a1[i] + 2.0 * a2[i] for i in range(NQA)], u)]
        return [q[i] + (a1[i] + 4.0 * a2[i] + a3[i]) * (1.0 / 6.0) for i in range(NQA)]

    def fd_agent(self, q, u):
        return np.array(self._rk3(list(q), list(u)), dtype=float)

    def rollout(self, u, x0):
        x = np.zeros((self.N + 1, self.n_q))
        x[0] = x0
        for k in range(self.N):
            for a in range(self.M):
                x[k + 1, NQA * a:NQA * (a + 1)] = self.fd_agent(x[k, NQA * a:NQA * (a + 1)],
                                                                u[self.uidx(a, k, 0):self.uidx(a, k, 0) + NUA])
        return x

    def linearize(self, x, u, order=2):
        """Dense joint A[N,nq,nq], B[N,nq,nu] and second derivatives T[N,nq,nq+nu,nq+nu] (joint variable
        order [x_k; u_k]), from second-order jets of the RK3 map (fAd/fBd/fEd/fFd/fGd, dynamics_models.py:128-144)."""
        N, M, nq, nu = self.N, self.M, self.n_q, self.n_u
        xq = x[:N].reshape(N * M, NQA)
        uu = np.stack([self._u_stage(u, k) for k in range(N)]).reshape(N * M, NUA)
        var = Jet.variables(np.hstack([xq, uu]), order=order)
        out = self._rk3(var[:NQA], var[NQA:])
        Jg = np.stack([j.g for j in out], axis=1).reshape(N, M, NQA, NQA + NUA)
        A = np.zeros((N, nq, nq))
        B = np.zeros((N, nq, nu))
        T = np.zeros((N, nq, nq + nu, nq + nu)) if order >= 2 else None
        if order >= 2:
            Jh = np.stack([j.h for j in out], axis=1).reshape(N, M, NQA, NQA + NUA, NQA + NUA)
        for a in range(M):
            sq, su = slice(NQA * a, NQA * (a + 1)), slice(NUA * a, NUA * (a + 1))
            A[:, sq, sq] = Jg[:, a, :, :NQA]
            B[:, sq, su] = Jg[:, a, :, NQA:]
            if order >= 2:
                idx = np.r_[NQA * a:NQA * (a + 1), nq + NUA * a:nq + NUA * (a + 1)]
                T[np.ix_(np.arange(N), np.arange(NQA * a, NQA * (a + 1)), idx, idx)] = Jh[:, a]
        return A, B, T

    def sensitivities(self, A, B):
        """f_Du_x (DGSQP.py:642-650): S[k] = d x_k / d u, columns agent-major."""
        S = np.zeros((self.N + 1, self.n_q, self.n))
        for k in range(self.N):
            S[k + 1] = A[k] @ S[k]
            S[k + 1][:, self.perm_inv[self.n_u * k:self.n_u * (k + 1)]] += B[k]
        return S

    # ------------------------------------------------------------------ costs
    def _state_cost(self, x_k, a):
        d = x_k[NQA * a:NQA * (a + 1)] - self.goals[a]
        return 0.5 * float(d @ (self.w_q * d))

    def costs(self, x, u, up):
        """f_J (DGSQP.py:889-893); stage k = 0..N-1: 1/2 w_u |u_k|^2 + state cost at x_k; terminal 10 x state cost."""
        N = self.N
        J = np.zeros(self.M)
        for a in range(self.M):
            ua = u[a * N * NUA:(a + 1) * N * NUA].reshape(N, NUA)
            J[a] = 0.5 * np.sum(self.w_u * ua ** 2) + sum(self._state_cost(x[k], a) for k in range(N)) \
                + self.term_scale * self._state_cost(x[N], a)
        return J

    def _cost_lx_lxx(self, x, a):
        N = self.N
        lx = np.zeros((N + 1, self.n_q))
        lxx = np.zeros((N + 1, self.n_q, self.n_q))
        sl = slice(NQA * a, NQA * (a + 1))
        for k in range(N + 1):
            sc = self.term_scale if k == N else 1.0
            lx[k, sl] = sc * self.w_q * (x[k, sl] - self.goals[a])
            lxx[k, sl, sl] = sc * np.diag(self.w_q)
        return lx, lxx

    # ------------------------------------------------------------- constraints
    def _lane_normal(self, a, j, px):
        brk, n_a, n_b, pt = self.lanes[a][j]
        return (n_b if px >= brk else n_a), pt

    def constraints(self, x, u, up):
        """f_Cxu (DGSQP.py:804-821,911)."""
        g = np.zeros(self.m)
        for r, (k, kind, a, b) in enumerate(self.rows):
            if kind == COLL:
                d = x[k, NQA * a:NQA * a + 2] - x[k, NQA * b:NQA * b + 2]
                g[r] = (self.obs_r[a] + self.obs_r[b]) ** 2 - d @ d
            elif kind == LANE:
                p = x[k, NQA * a:NQA * a + 2]
                nrm, pt = self._lane_normal(a, b, p[0])
                g[r] = nrm @ (p - (pt - self.lane_r * nrm))
            elif kind == IN_UB:
                g[r] = u[self.uidx(a, k, b)] - self.u_ub[b]
            elif kind == IN_LB:
                g[r] = self.u_lb[b] - u[self.uidx(a, k, b)]
            elif kind == ST_UB:
                g[r] = x[k, NQA * a + b] - self.v_ub
            elif kind == ST_LB:
                g[r] = self.v_lb - x[k, NQA * a + b]
        return g

    def constraint_jacobian(self, x, S):
        """f_Du_C (DGSQP.py:824-826,918): G = dC/du + dC/dx Du_x.  pw_const has zero derivative."""
        G = np.zeros((self.m, self.n))
        for r, (k, kind, a, b) in enumerate(self.rows):
            if kind == COLL:
                d = x[k, NQA * a:NQA * a + 2] - x[k, NQA * b:NQA * b + 2]
                G[r] = -2.0 * d @ (S[k][NQA * a:NQA * a + 2] - S[k][NQA * b:NQA * b + 2])
            elif kind == LANE:
                nrm, _ = self._lane_normal(a, b, x[k, NQA * a])
                G[r] = nrm @ S[k][NQA * a:NQA * a + 2]
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
            elif kind == LANE:
                nrm, _ = self._lane_normal(a, b, x[k, NQA * a])
                lx[k, NQA * a:NQA * a + 2] += l[r] * nrm
            elif kind == ST_UB:
                lx[k, NQA * a + b] += l[r]
            elif kind == ST_LB:
                lx[k, NQA * a + b] -= l[r]
        return lx, lxx

    # ------------------------------------------------------------ gradient q
    def cost_gradient(self, x, u, up, S):
        """f_q (DGSQP.py:673-676,898-899): q = [grad_{u^a} J^a]_a."""
        N = self.N
        q = np.zeros(self.n)
        for a in range(self.M):
            sl = slice(a * N * NUA, (a + 1) * N * NUA)
            lx, _ = self._cost_lx_lxx(x, a)
            ga = np.tile(self.w_u, N) * u[sl]
            for k in range(1, N + 1):
                ga = ga + lx[k] @ S[k][:, sl]
            q[sl] = ga
        return q

    def sum_cost_gradient(self, x, u, up, S):
        """grad_u sum_a J^a (the `dobj` of the v2 merit 'sum_obj_l1', DGSQP_v2.py:1150-1151)."""
        N = self.N
        lx = sum(self._cost_lx_lxx(x, f)[0] for f in range(self.M))
        qs = np.tile(self.w_u, N * self.M) * u
        for k in range(1, N + 1):
            qs = qs + lx[k] @ S[k]
        return qs

    # ------------------------------------------------------------- Hessian Q
    def _dp_hessian(self, A, B, T, lx, lxx, luu):
        """Hessian wrt the stage-major joint input sequence of Phi(u) = sum_k [l_k(x_k) + 1/2 u_k' luu u_k]
        + l_N(x_N) along x_{k+1} = f(x_k, u_k) by the backward recursion of DGSQP.py:679-727 / :829-877
        (Dx_Q -> p, Dxx_Q -> V, Dxu_Q -> W rows, Duu_Q -> D)."""
        N, nq, nu = self.N, self.n_q, self.n_u
        p, V = lx[N].copy(), lxx[N].copy()
        W = np.zeros((0, nq))
        D = np.zeros((0, 0))
        for k in range(N - 1, -1, -1):
            Tp = np.tensordot(p, T[k], axes=(0, 0))            # sum_i p_i Hess(f_i), (nq+nu, nq+nu)
            E, Gm, F = Tp[:nq, :nq], Tp[nq:, :nq], Tp[nq:, nq:]
            D11 = luu + B[k].T @ V @ B[k] + F
            D21 = W @ B[k]
            D = np.block([[D11, D21.T], [D21, D]]) if W.shape[0] else D11
            W = np.vstack([B[k].T @ V @ A[k] + Gm, W @ A[k]])
            V = lxx[k] + A[k].T @ V @ A[k] + E
            p = lx[k] + p @ A[k]
        return D[np.ix_(self.perm, self.perm)]

    def hessian(self, x, u, l, A, B, T):
        """f_Q (DGSQP.py:920-934): row block a = grad_{u^a} grad_u (J^a + l'C)."""
        N, nu = self.N, self.n_u
        Q = np.zeros((self.n, self.n))
        for a in range(self.M):
            lx, lxx = self._cost_lx_lxx(x, a)
            luu = np.zeros((nu, nu))
            luu[NUA * a:NUA * (a + 1), NUA * a:NUA * (a + 1)] = np.diag(self.w_u)
            sl = slice(a * N * NUA, (a + 1) * N * NUA)
            Q[sl] = self._dp_hessian(A, B, T, lx, lxx, luu)[sl]
        lx, lxx = self._constraint_lx_lxx(x, l)
        Q += self._dp_hessian(A, B, T, lx, lxx, np.zeros((nu, nu)))
        return Q

    # ---------------------------------------------------------------- evaluate
    def evaluate(self, u, l, x0, up, hessian=True):
        """_evaluate (DGSQP.py:509-533).  Returns (Q, q, G, g, x) or (q, G, g, x)."""
        x = self.rollout(u, x0)
        A, B, T = self.linearize(x, u, order=2 if hessian else 1)
        S = self.sensitivities(A, B)
        g = self.constraints(x, u, up)
        G = self.constraint_jacobian(x, S)
        q = self.cost_gradient(x, u, up, S)
        # kept for the v2 merit 'sum_obj_l1' (sum of the costs and its gradient at the evaluated point)
        self.last_obj, self.last_qs = float(np.sum(self.costs(x, u, up))), self.sum_cost_gradient(x, u, up, S)
        if hessian:
            return self.hessian(x, u, l, A, B, T), q, G, g, x
        return q, G, g, x


def sample_merge(num, seed=1, game=None):
    """Initial conditions of scripts/DGSQP_merge_monte_carlo.py:428-495 (rng = default_rng(seed), same draw
    order).  The script's collision pre-check rolls the zero-input warm start forward; it never initialises
    ``car3_q_ws[0]`` (:486-488), so car 3 is checked from the origin state -- mirrored here.
    Returns x0[num, 12]; the warm start of every instance is u = 0 (DGSQP.py:179)."""
    game = MergeGame() if game is None else game
    rng = np.random.default_rng(seed)
    th = math.pi / 12
    lanes = merge_lanes()
    x5_0, x7_0 = 1.5, lanes[2][1][0]
    out = []
    while len(out) < num:
        cars = []
        for x_nom in (0.0, 0.5):
            x = x_nom + 0.5 * rng.random() - 0.25
            y = 0.15 + 0.1 * rng.random() - 0.05
            v = 0.3 * (1 + 0.06 * rng.random() - 0.03)
            p = 0.0 + (5 * rng.random() - 2.5) * np.pi / 180
            cars.append([x, y, v, p])
        x_nom, y_nom = 0.25, -((x7_0 + x5_0) / 2 - 0.25) * np.tan(th)
        s = 0.5 * rng.random() - 0.25
        ey = 0.1 * rng.random() - 0.05
        x = x_nom + s * np.cos(th) - ey * np.sin(th)
        y = y_nom + s * np.sin(th) + ey * np.cos(th)
        v = 0.3 * (1 + 0.06 * rng.random() - 0.03)
        p = np.pi / 12 + (5 * rng.random() - 2.5) * np.pi / 180
        cars.append([x, y, v, p])
        # collision pre-check on the zero-input rollout (car 3 from the zero state, as in the script)
        trajs = []
        for a, q0 in enumerate([cars[0], cars[1], [0.0, 0.0, 0.0, 0.0]]):
            q = np.zeros((game.N + 1, NQA))
            q[0] = q0
            for k in range(game.N):
                q[k + 1] = game.fd_agent(q[k], [0.0, 0.0])
            trajs.append(q)
        hit = False
        for i in range(3):
            for j in range(i + 1, 3):
                d = np.linalg.norm(trajs[i][:, :2] - trajs[j][:, :2], axis=1)
                if np.any(d < game.obs_r[i] + game.obs_r[j]):
                    hit = True
        if hit:
            continue
        out.append(np.concatenate(cars))
    return np.array(out)
