"""Front end for the reference's own constructor arguments.

The reference builds its solver as ``DGSQP(joint_dynamics, costs, agent_constraints, shared_constraints, bounds,
params)`` (``DGSQP/solvers/DGSQP.py:26-34``) where ``costs`` / ``*_constraints`` are lists of ``ca.Function`` objects the
driver scripts assemble (``scripts/DGSQP_ALGAMES_monte_carlo_chicane.py:222-339``).  CasADi graphs cannot be evaluated on
the GPU (and CasADi is not installable offline); the device solver implements the game FAMILY those scripts instantiate
-- kinematic bicycles in Frenet coordinates with quadratic input / input-rate stage costs, progress + competition terminal
costs, input-rate rows, pairwise collision rows and box bounds -- with the numbers as data.

:func:`game_from_reference_args` takes the same six arguments with plain Python callables in place of the
``ca.Function`` objects (same signatures: stage ``f(q, u, um)``, terminal ``f(q)``, ``None`` where the scripts pass
``None``), IDENTIFIES the numbers of the family from them -- every function is probed at random points, the family's
parameters are fitted (the stage cost is linear in its weights, the rows are affine) and the fit is verified on fresh
points to 1e-9 -- and returns the :class:`~dgsqp_b200.games.RacingGame` record the C ABI consumes.  Anything that is not a
member of the family (an extra cost term, a different row, agents with different weights) raises
``NotImplementedError`` naming the function that did not fit; nothing is silently approximated.
``dgsqp_b200.DGSQP`` accepts the six-argument form directly and routes it through here.
"""
import math
from typing import Callable, List, Optional, Sequence

import numpy as np

from .dynamics import CasadiDecoupledMultiAgentDynamicsModel, CasadiKinematicBicycleCombined, CasadiKinematicUnicycle
from .games import NQA, NUA, RacingGame

_TOL = 1e-9
_IDX = dict(x=0, y=1, v=2, e_psi=3, s=4, x_tran=5)


class UnsupportedGameError(NotImplementedError):
    pass


def _snap(v: float) -> float:
    """Identified numbers carry the rounding of the fit (1.0000000000000002 for a weight of 1): values within 1e-12 of a
    10-digit decimal are returned as that decimal."""
    r = round(float(v), 10)
    return r if abs(r - float(v)) <= 1e-12 * max(1.0, abs(r)) else float(v)


def _fit(fn: Callable, feats: Callable, nfeat: int, sample: Callable, what: str, rng) -> np.ndarray:
    """Least-squares coefficients c with fn(z) = feats(z) . c, verified on fresh samples."""
    Z = [sample(rng) for _ in range(4 * nfeat + 8)]
    A = np.array([feats(*z) for z in Z])
    b = np.array([float(np.asarray(fn(*z)).squeeze()) for z in Z])
    c, *_ = np.linalg.lstsq(A, b, rcond=None)
    for _ in range(16):
        z = sample(rng)
        got, want = float(np.asarray(fn(*z)).squeeze()), float(np.dot(feats(*z), c))
        if abs(got - want) > _TOL * max(1.0, abs(got)):
            raise UnsupportedGameError(f"{what} is not a member of the supported game family "
                                       f"(fit residual {abs(got - want):.2e} at a probe point)")
    return c


def _rows(fn: Callable, args, what: str) -> np.ndarray:
    v = np.atleast_1d(np.asarray(fn(*args), dtype=np.float64)).ravel()
    if not np.all(np.isfinite(v)):
        raise UnsupportedGameError(f"{what} returned non-finite values at a probe point")
    return v


def game_from_reference_args(joint_dynamics, costs: Sequence[Sequence[Callable]],
                             agent_constraints: Sequence[Sequence[Optional[Callable]]],
                             shared_constraints: Sequence[Optional[Callable]], bounds: dict, params=None, seed: int = 0):
    """The game record of the reference-style arguments (see the module docstring).  ``params`` only supplies ``N`` / ``dt``
    cross-checks (``DGSQPParams.N``, ``.dt``)."""
    if not isinstance(joint_dynamics, CasadiDecoupledMultiAgentDynamicsModel):
        raise UnsupportedGameError("joint_dynamics must be a CasadiDecoupledMultiAgentDynamicsModel of kinematic bicycles "
                                   "(dgsqp_b200.dynamics); coupled joint models are not part of the device game family")
    models = joint_dynamics.dynamics_models
    M = len(models)
    if any(isinstance(m, CasadiKinematicUnicycle) for m in models):
        raise UnsupportedGameError("unicycle agents: use dgsqp_b200.merge_game() -- the merge scenario's lane rows are "
                                   "piecewise-constant half-planes (scripts/DGSQP_merge_monte_carlo.py:66-74) that a "
                                   "point-probing front end cannot identify")
    if not all(isinstance(m, CasadiKinematicBicycleCombined) for m in models):
        raise UnsupportedGameError("every agent must be a CasadiKinematicBicycleCombined")
    m0 = models[0]
    for m in models[1:]:
        if (m.L_f, m.L_r, m.m, m.c_dr, m.c_da, m.c_s, m.dt) != (m0.L_f, m0.L_r, m0.m, m0.c_dr, m0.c_da, m0.c_s, m0.dt) \
                or m.track is not m0.track:
            raise UnsupportedGameError("agents with different vehicle parameters or tracks")
    if m0.track is None:
        raise UnsupportedGameError("the bicycle models carry no track")
    if len(costs) != M or len(agent_constraints) != M:
        raise ValueError("Number of agents: %i, but %i cost / %i constraint lists were provided" % (M, len(costs), len(agent_constraints)))
    N = len(costs[0]) - 1
    if params is not None and getattr(params, "N", N) != N:
        raise ValueError(f"params.N = {params.N} but the cost lists describe N = {N}")
    dt = m0.dt
    rng = np.random.default_rng(seed)
    nq, nu = NQA * M, NUA
    sq = lambda r: r.normal(size=nq) * np.tile([2.0, 2.0, 1.0, 0.3, 3.0, 0.5], M)
    su = lambda r: r.normal(size=nu)

    # ---- stage costs: 1/2 sum_c w_u[c] u_c^2 + 1/2 sum_c w_du[c] (u_c - um_c)^2, identical over stages and agents
    feats = lambda q, u, um: np.array([0.5 * u[0] ** 2, 0.5 * u[1] ** 2, 0.5 * (u[0] - um[0]) ** 2, 0.5 * (u[1] - um[1]) ** 2])
    w_all = []
    for a in range(M):
        if len(costs[a]) != N + 1:
            raise ValueError("cost lists of different lengths")
        for k in sorted({0, N // 2, N - 1}):
            w_all.append(_fit(costs[a][k], feats, 4, lambda r: (sq(r), su(r), su(r)), f"stage cost {k} of agent {a}", rng))
    w = np.array(w_all)
    if np.abs(w - w[0]).max() > 1e-9 * max(1.0, np.abs(w[0]).max()):
        raise UnsupportedGameError("stage-cost weights differ between agents or stages (the device game carries one set)")
    # ---- terminal costs: -c_prog s_a + sum_{b != a} c_comp atan(s_b - s_a)
    c_all = []
    for a in range(M):
        tf = lambda q, a=a: np.array([-q[a * NQA + _IDX["s"]],
                                      sum(math.atan(q[b * NQA + _IDX["s"]] - q[a * NQA + _IDX["s"]]) for b in range(M) if b != a)])
        c_all.append(_fit(costs[a][N], tf, 2, lambda r: (sq(r),), f"terminal cost of agent {a}", rng))
    cw = np.array(c_all)
    if np.abs(cw - cw[0]).max() > 1e-9 * max(1.0, np.abs(cw[0]).max()):
        raise UnsupportedGameError("terminal-cost weights differ between agents")
    # ---- agent rows: (du_a - dt r_ub_a, dt r_lb_a - du_a, du_s - dt r_ub_s, dt r_lb_s - du_s), no terminal rows
    rate = None
    for a in range(M):
        if len(agent_constraints[a]) != N + 1 or agent_constraints[a][N] is not None:
            raise UnsupportedGameError(f"agent {a}: terminal agent constraints are not part of the family")
        for k in sorted({0, N - 1}):
            fn = agent_constraints[a][k]
            q, u, um = sq(rng), su(rng), su(rng)
            c0 = _rows(fn, (q, np.zeros(nu), np.zeros(nu)), f"agent constraint {k} of agent {a}")
            if c0.size != 4:
                raise UnsupportedGameError(f"agent constraint {k} of agent {a} has {c0.size} rows, expected the 4 rate rows")
            du = u - um
            want = c0 + np.array([du[0], -du[0], du[1], -du[1]])
            if np.abs(_rows(fn, (q, u, um), "agent constraint") - want).max() > _TOL * max(1.0, np.abs(want).max()):
                raise UnsupportedGameError(f"agent constraint {k} of agent {a} is not the input-rate box of the family")
            r = (-c0[0] / dt, c0[1] / dt, -c0[2] / dt, c0[3] / dt)          # r_ub_a, r_lb_a, r_ub_s, r_lb_s
            if rate is not None and np.abs(np.array(r) - np.array(rate)).max() > 1e-9:
                raise UnsupportedGameError("input-rate limits differ between agents or stages")
            rate = r
    # ---- shared rows: none at stage 0, (r_a + r_b)^2 - |p_a - p_b|^2 per agent pair at stages 1..N
    if len(shared_constraints) != N + 1 or shared_constraints[0] is not None:
        raise UnsupportedGameError("shared constraints: expected None at stage 0 and the collision rows at stages 1..N")
    pairs = [(a, b) for a in range(M) for b in range(a + 1, M)]
    d2 = None
    for k in sorted({1, N}):
        fn = shared_constraints[k]
        q = sq(rng)
        args = (q,) if k == N else (q, np.zeros(NUA * M), np.zeros(NUA * M))
        v = _rows(fn, args, f"shared constraint {k}")
        if v.size != len(pairs):
            raise UnsupportedGameError(f"shared constraint {k} has {v.size} rows, expected {len(pairs)} collision rows")
        dist2 = np.array([(q[a * NQA] - q[b * NQA]) ** 2 + (q[a * NQA + 1] - q[b * NQA + 1]) ** 2 for a, b in pairs])
        dk = v + dist2
        q2 = sq(rng)
        args2 = (q2,) if k == N else (q2, np.zeros(NUA * M), np.zeros(NUA * M))
        dist2b = np.array([(q2[a * NQA] - q2[b * NQA]) ** 2 + (q2[a * NQA + 1] - q2[b * NQA + 1]) ** 2 for a, b in pairs])
        if np.abs(_rows(fn, args2, "shared constraint") + dist2b - dk).max() > _TOL * max(1.0, np.abs(dk).max()) or np.any(dk <= 0):
            raise UnsupportedGameError(f"shared constraint {k} is not the pairwise collision row of the family")
        if d2 is not None and np.abs(dk - d2).max() > 1e-9:
            raise UnsupportedGameError("collision radii differ between stages")
        d2 = dk
    dsum = np.sqrt(d2)                                                    # r_a + r_b per pair
    if M == 2:
        obs_r = [dsum[0] / 2, dsum[0] / 2]
    else:
        A = np.zeros((len(pairs), M))
        for i, (a, b) in enumerate(pairs):
            A[i, a] = A[i, b] = 1.0
        obs_r, *_ = np.linalg.lstsq(A, dsum, rcond=None)
        if np.abs(A @ obs_r - dsum).max() > 1e-9:
            raise UnsupportedGameError("collision distances are not sums of per-agent radii")
        obs_r = list(obs_r)
    # ---- box bounds: per-agent VehicleState records; inputs and the lateral offset are the bounded entries of the family
    ub, lb = bounds["ub"], bounds["lb"]
    if len(ub) != M or len(lb) != M:
        raise ValueError("bounds: one VehicleState per agent expected")
    u_ub, u_lb, hw = (ub[0].u.u_a, ub[0].u.u_steer), (lb[0].u.u_a, lb[0].u.u_steer), ub[0].p.x_tran
    for a in range(M):
        if (ub[a].u.u_a, ub[a].u.u_steer) != u_ub or (lb[a].u.u_a, lb[a].u.u_steer) != u_lb \
                or ub[a].p.x_tran != hw or lb[a].p.x_tran != -hw:
            raise UnsupportedGameError("box bounds differ between agents or are not symmetric in the lateral offset")
        for name, val in (("x.x", ub[a].x.x), ("x.y", ub[a].x.y), ("p.s", ub[a].p.s), ("p.e_psi", ub[a].p.e_psi), ("v.v_long", ub[a].v.v_long)):
            if np.isfinite(val):
                raise UnsupportedGameError(f"a finite bound on {name} is not part of the family (only inputs and p.x_tran)")
    return RacingGame(track=m0.track, M=M, N=N, dt=dt, L_f=m0.L_f, L_r=m0.L_r, c_dr=m0.c_dr, c_da=m0.c_da, c_s=m0.c_s,
                      mass=m0.m, input_weight=(_snap(w[0][0]), _snap(w[0][1])), rate_weight=(_snap(w[0][2]), _snap(w[0][3])),
                      comp_weights=(_snap(cw[0][0]), _snap(cw[0][1])), u_ub=tuple(map(float, u_ub)), u_lb=tuple(map(float, u_lb)),
                      rate_ub=(_snap(rate[0]), _snap(rate[2])), rate_lb=(_snap(rate[1]), _snap(rate[3])), half_width=float(hw),
                      obs_r=[_snap(r) for r in obs_r], name="reference_args")
