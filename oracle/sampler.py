"""Monte-Carlo instance sampler with PID warm start.  Oracle-only restatement.

Follows ``scripts/DGSQP_ALGAMES_monte_carlo_chicane.py:381-485`` (2-agent head-to-head,
also the curve script) and ``scripts/DGSQP_monte_carlo_agents.py:257-325`` (M independent
agents); the PID controllers follow ``DGSQP/solvers/PID.py:74-138,187-238`` and the one-step
simulation ``CasadiDynamicsModel.step`` (``DGSQP/dynamics/dynamics_models.py:161-186``:
SciPy ``solve_ivp`` RK45 with default tolerances on the continuous model, then
``local_to_global`` on the result).  The reference scripts for chicane/agents are unseeded;
this restatement takes the seed as an argument and draws in the script's order.
"""
import numpy as np
from scipy.integrate import solve_ivp


class _PID:
    def __init__(self, dt, Kp, Ki, x_ref, u_max, u_min, du_max, du_min):
        self.dt, self.Kp, self.Ki, self.Kd = dt, Kp, Ki, 0.0
        self.int_e_max, self.int_e_min = 100, -100
        self.u_max, self.u_min, self.du_max, self.du_min = u_max, u_min, du_max, du_min
        self.x_ref, self.u_ref, self.u_prev = x_ref, 0.0, 0
        self.e = self.de = self.ei = 0

    def solve(self, x):
        u_prev = self.u_prev
        e_t = x - self.x_ref
        de_t = (e_t - self.e) / self.dt
        ei_t = self.ei + e_t * self.dt
        ei_t = min(max(ei_t, self.int_e_min), self.int_e_max)
        u = -(self.Kp * e_t + self.Ki * ei_t + self.Kd * de_t) + self.u_ref
        du = u - u_prev
        du = np.minimum(du, self.du_max)
        du = np.maximum(du, self.du_min)
        u = du + u_prev
        u = np.minimum(u, self.u_max)
        u = np.maximum(u, self.u_min)
        self.e, self.de, self.ei = e_t, de_t, ei_t
        self.u_prev = u
        return u


class _LaneFollower:
    """PIDLaneFollower (PID.py:187-238): lat_ref = steer x_ref, steer PID then tracks 0."""

    def __init__(self, dt, steer, speed):
        self.steer, self.speed = steer, speed
        self.lat_ref = steer.x_ref
        self.steer.x_ref = 0
        self.steer.ei = 0
        self.steer.e = 0

    def step(self, v_long, x_tran, e_psi):
        u_a = self.speed.solve(v_long)
        u_s = self.steer.solve(5.0 * (x_tran - self.lat_ref) + 1.0 * e_psi)
        return float(u_a), float(u_s)


def pid_rollout(game, s0, xtran0, v0):
    """N steps of PID + RK45 for one agent starting at (s0, x_tran0, v0), e_psi=0.
    Returns q_ws[N+1,6] (with the script's s-1e-6) and u_ws[N,2]."""
    N, dt, track = game.N, game.dt, game.track
    steer = _PID(dt, 1.0, 0.005, xtran0, game.u_ub[1], game.u_lb[1], game.rate_ub[1], game.rate_lb[1])
    speed = _PID(dt, 1.0, 0.0, v0, game.u_ub[0], game.u_lb[0], game.rate_ub[0], game.rate_lb[0])
    pid = _LaneFollower(dt, steer, speed)
    x, y, _ = track.local_to_global((s0, xtran0, 0.0))
    q = np.array([x, y, v0, 0.0, s0, xtran0])
    qs, us = [q.copy()], []
    for _ in range(N):
        u = np.array(pid.step(q[2], q[5], q[3]))
        sol = solve_ivp(lambda t, z: game._fc_scalar(z, u), (0, dt), q, t_eval=[dt])
        q = sol.y.squeeze().copy()
        gx, gy, _ = track.local_to_global((q[4], q[5], q[3]))
        q[0], q[1] = gx, gy
        qs.append(q.copy())
        us.append(u)
    qs = np.array(qs)
    q_ws = qs.copy()
    q_ws[:, 4] -= 1e-6
    return qs[0], q_ws, np.array(us)


def _collides(trajs, radii):
    M = len(trajs)
    for i in range(M):
        for j in range(i + 1, M):
            d = np.linalg.norm(trajs[i][:, :2] - trajs[j][:, :2], axis=1)
            if np.any(d < radii[i] + radii[j]):
                return True
    return False


def sample_head_to_head(game, rng, ego_r=None, tar_r=None):
    """One accepted chicane/curve instance: returns x0[12], u_ws agent-major [n]."""
    first_seg_len = game.track.cl_segs[0, 0]
    hw = game.half_width
    obs_d = (game.obs_r[0] if ego_r is None else ego_r) + (game.obs_r[1] if tar_r is None else tar_r)
    while True:
        ego_s = max(0.1, rng.random() * first_seg_len)
        ego_xt = rng.random() * hw * 2 - hw
        ego_v = rng.random() + 2
        d = 2 * np.pi * rng.random()
        tar_s = ego_s + 1.2 * obs_d * np.cos(d)
        if tar_s < 0:
            continue
        tar_xt = ego_xt + 1.2 * obs_d * np.sin(d)
        if np.abs(tar_xt) > hw:
            continue
        tar_v = rng.random() + 2
        e0, eq, eu = pid_rollout(game, ego_s, ego_xt, ego_v)
        t0, tq, tu = pid_rollout(game, tar_s, tar_xt, tar_v)
        if not _collides([eq, tq], [obs_d / 2, obs_d / 2]):
            break
    x0 = np.concatenate([e0, t0])
    return x0, np.concatenate([eu.ravel(), tu.ravel()])


def sample_agents(game, rng):
    """One accepted M-agent instance (DGSQP_monte_carlo_agents.py:257-308)."""
    first_seg_len = game.track.cl_segs[0, 0]
    hw = game.half_width
    while True:
        q0s, qws, uws = [], [], []
        for _ in range(game.M):
            s = max(0.1, rng.random() * first_seg_len)
            xt = rng.random() * hw * 2 - hw
            v = rng.random() + 2
            q0, qw, uw = pid_rollout(game, s, xt, v)
            q0s.append(q0)
            qws.append(qw)
            uws.append(uw)
        if not _collides(qws, game.obs_r):
            break
    return np.concatenate(q0s), np.concatenate([u.ravel() for u in uws])
