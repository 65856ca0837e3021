"""Independent float64 torch.autograd statement of the racing game's batch functions
J^a(u), C(u) (DGSQP.py:889-915).  Used only to pin the oracle's derivatives: it is the same
check as the reference's f_Du_L / f_Duu_L (DGSQP.py:937-941) against its DP Hessian f_Q."""
import math
import torch

torch.set_default_dtype(torch.float64)


def _track_terms(track, s):
    L = track.track_length
    sb = torch.fmod(torch.fmod(s, L) + L, L)
    idx = int(sum(sb.item() >= b for b in track.curv_breaks))
    kap = float(track.curv_vals[idx])
    psit = float(track.cum_ang[idx]) + float(track.slopes[idx]) * (sb - float(track.cum_len[idx]))
    return kap, psit


def rollout(game, u, x0):
    """u: agent-major torch vector (requires_grad), x0: list/array. Returns list of joint states."""
    M, N, dt = game.M, game.N, game.dt
    x = [torch.tensor(x0, dtype=torch.float64)]
    L = game.L_f + game.L_r
    for k in range(N):
        nxt = []
        for a in range(M):
            q = x[k][6 * a:6 * a + 6]
            acc, delta = u[a * N * 2 + 2 * k], u[a * N * 2 + 2 * k + 1]
            v, epsi, s, ey = q[2], q[3], q[4], q[5]
            beta = torch.atan2(torch.tan(delta) * game.L_r, torch.tensor(L))
            psidot = v / game.L_r * torch.sin(beta)
            absv = v if v.item() > 0 else -v
            F = -game.c_da * v - game.c_dr * v * absv - game.c_s * psidot ** 2
            kap, psit = _track_terms(game.track, s)
            den = 1 - ey * kap
            f = torch.stack([v * torch.cos(beta + psit + epsi), v * torch.sin(beta + psit + epsi),
                             acc + F / game.mass, psidot - kap * v * torch.cos(beta + epsi) / den,
                             v * torch.cos(beta + epsi) / den, v * torch.sin(beta + epsi)])
            nxt.append(q + dt * f)
        x.append(torch.cat(nxt))
    return x


def costs(game, u, x, up):
    M, N = game.M, game.N
    J = []
    for a in range(M):
        ua = u[a * N * 2:(a + 1) * N * 2].reshape(N, 2)
        um = torch.cat([torch.tensor(up[2 * a:2 * a + 2]).reshape(1, 2), ua[:-1]])
        wu, wdu = torch.tensor(game.w_u), torch.tensor(game.w_du)
        Ja = 0.5 * (wu * ua ** 2).sum() + 0.5 * (wdu * (ua - um) ** 2).sum()
        sN = x[N][6 * a + 4]
        Ja = Ja - game.c_prog * sN
        for b in range(M):
            if b != a:
                Ja = Ja + game.c_comp * torch.atan(x[N][6 * b + 4] - sN)
        J.append(Ja)
    return J


def constraints(game, u, x, up):
    from oracle.racing_game import COLL, RATE, IN_UB, IN_LB, ST_UB, ST_LB
    N = game.N
    out = []
    for (k, kind, a, b) in game.rows:
        if kind == COLL:
            d = x[k][6 * a:6 * a + 2] - x[k][6 * b:6 * b + 2]
            out.append((game.obs_r[a] + game.obs_r[b]) ** 2 - (d * d).sum())
        elif kind == RATE:
            c = b // 2
            uk = u[a * N * 2 + 2 * k + c]
            um = torch.tensor(up[2 * a + c]) if k == 0 else u[a * N * 2 + 2 * (k - 1) + c]
            du = uk - um
            out.append(du - game.dt * game.rate_ub[c] if b % 2 == 0 else game.dt * game.rate_lb[c] - du)
        elif kind == IN_UB:
            out.append(u[a * N * 2 + 2 * k + b] - game.u_ub[b])
        elif kind == IN_LB:
            out.append(game.u_lb[b] - u[a * N * 2 + 2 * k + b])
        elif kind == ST_UB:
            out.append(x[k][6 * a + b] - game.half_width)
        elif kind == ST_LB:
            out.append(-game.half_width - x[k][6 * a + b])
    return torch.stack(out)


def evaluate_autograd(game, u_np, l_np, x0, up):
    """Returns (Q, q, G, g) by direct differentiation of the batch functions."""
    import numpy as np
    N, M, n = game.N, game.M, game.n
    l = torch.tensor(l_np)

    def gradL(uvec):
        x = rollout(game, uvec, x0)
        J = costs(game, uvec, x, up)
        C = constraints(game, uvec, x, up)
        rows = []
        for a in range(M):
            La = J[a] + (l * C).sum()
            ga = torch.autograd.grad(La, uvec, create_graph=True)[0]
            rows.append(ga[a * N * 2:(a + 1) * N * 2])
        return torch.cat(rows)

    u = torch.tensor(u_np, requires_grad=True)
    Q = torch.autograd.functional.jacobian(gradL, u).numpy()
    x = rollout(game, u, x0)
    J = costs(game, u, x, up)
    C = constraints(game, u, x, up)
    q = np.concatenate([torch.autograd.grad(J[a], u, retain_graph=True)[0].numpy()[a * N * 2:(a + 1) * N * 2]
                        for a in range(M)])
    G = torch.autograd.functional.jacobian(lambda uu: constraints(game, uu, rollout(game, uu, x0), up), u).numpy()
    return Q, q, G, C.detach().numpy()
