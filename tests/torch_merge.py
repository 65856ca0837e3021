"""Independent float64 torch.autograd statement of the merge game's batch functions J^a(u), C(u)
(scripts/DGSQP_merge_monte_carlo.py:253-398 through DGSQP.py:889-915).  Used only to pin the oracle's derivatives
(oracle/merge_game.py): the same check as the reference's f_Du_L / f_Duu_L (DGSQP.py:937-941) against its DP Hessian."""
import numpy as np
import torch

torch.set_default_dtype(torch.float64)


def _fc(game, q, u):
    return torch.stack([q[2] * torch.cos(q[3]), q[2] * torch.sin(q[3]), u[0] / game.mass, u[1]])


def rollout(game, u, x0):
    M, N, h = game.M, game.N, game.dt
    x = [torch.tensor(x0, dtype=torch.float64)]
    for k in range(N):
        nxt = []
        for a in range(M):
            q = x[k][4 * a:4 * a + 4]
            uk = u[a * N * 2 + 2 * k:a * N * 2 + 2 * k + 2]
            a1 = h * _fc(game, q, uk)
            a2 = h * _fc(game, q + a1 / 2, uk)
            a3 = h * _fc(game, q - a1 + 2 * a2, uk)
            nxt.append(q + (a1 + 4 * a2 + a3) / 6)
        x.append(torch.cat(nxt))
    return x


def costs(game, u, x):
    M, N = game.M, game.N
    wq, wu = torch.tensor(game.w_q), torch.tensor(game.w_u)
    J = []
    for a in range(M):
        ua = u[a * N * 2:(a + 1) * N * 2].reshape(N, 2)
        goal = torch.tensor(game.goals[a])
        Ja = 0.5 * (wu * ua ** 2).sum()
        for k in range(N + 1):
            d = x[k][4 * a:4 * a + 4] - goal
            Ja = Ja + (game.term_scale if k == N else 1.0) * 0.5 * (wq * d * d).sum()
        J.append(Ja)
    return J


def constraints(game, u, x):
    from oracle.merge_game import COLL, LANE, IN_UB, IN_LB, ST_UB, ST_LB
    N = game.N
    out = []
    for (k, kind, a, b) in game.rows:
        if kind == COLL:
            d = x[k][4 * a:4 * a + 2] - x[k][4 * b:4 * b + 2]
            out.append((game.obs_r[a] + game.obs_r[b]) ** 2 - (d * d).sum())
        elif kind == LANE:
            p = x[k][4 * a:4 * a + 2]
            brk, n_a, n_b, pt = game.lanes[a][b]
            nrm = torch.tensor(np.asarray(n_b if p[0].item() >= brk else n_a, dtype=float))   # pw_const: no gradient
            out.append((nrm * (p - (torch.tensor(np.asarray(pt, dtype=float)) - game.lane_r * nrm))).sum())
        elif kind == IN_UB:
            out.append(u[a * N * 2 + 2 * k + b] - game.u_ub[b])
        elif kind == IN_LB:
            out.append(game.u_lb[b] - u[a * N * 2 + 2 * k + b])
        elif kind == ST_UB:
            out.append(x[k][4 * a + b] - game.v_ub)
        elif kind == ST_LB:
            out.append(game.v_lb - x[k][4 * a + b])
    return torch.stack(out)


def evaluate_autograd(game, u_np, l_np, x0):
    """Returns (Q, q, G, g) by direct differentiation of the batch functions."""
    N, M = game.N, game.M
    l = torch.tensor(l_np)

    def gradL(uvec):
        x = rollout(game, uvec, x0)
        J = costs(game, uvec, x)
        C = constraints(game, uvec, x)
        rows = []
        for a in range(M):
            ga = torch.autograd.grad(J[a] + (l * C).sum(), uvec, create_graph=True)[0]
            rows.append(ga[a * N * 2:(a + 1) * N * 2])
        return torch.cat(rows)

    u = torch.tensor(u_np, requires_grad=True)
    Q = torch.autograd.functional.jacobian(gradL, u).numpy()
    x = rollout(game, u, x0)
    J = costs(game, u, x)
    C = constraints(game, u, x)
    q = np.concatenate([torch.autograd.grad(J[a], u, retain_graph=True)[0].numpy()[a * N * 2:(a + 1) * N * 2]
                        for a in range(M)])
    G = torch.autograd.functional.jacobian(lambda uu: constraints(game, uu, rollout(game, uu, x0)), u).numpy()
    return Q, q, G, C.detach().numpy()
