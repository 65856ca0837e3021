"""Pins the oracle's solver layer: exact QP (KKT conditions + an independent SLSQP solve), nearestPD, merit
directional derivative by finite differences, and full solves against the committed golden values."""
import json
import pathlib

import numpy as np
import pytest
import scipy.optimize as so

from oracle.qp import solve_qp_gi, kkt_residuals, QPFailure
from oracle.dgsqp_v1 import OracleDGSQP, nearest_pd
from oracle.racing_game import RacingGame
from oracle.track import chicane_track

GOLDEN = pathlib.Path(__file__).parent / "golden"


@pytest.mark.parametrize("seed", range(6))
def test_qp_kkt_and_independent_solver(seed):
    rng = np.random.default_rng(seed)
    n, m = 12, 30
    A = rng.normal(size=(n, n))
    H = A @ A.T + 0.05 * np.eye(n)
    q = rng.normal(size=n) * 3
    G = rng.normal(size=(m, n))
    g = -np.abs(rng.normal(size=m)) * 0.5
    x, lam = solve_qp_gi(H, q, G, g)
    r = kkt_residuals(H, q, G, g, x, lam)
    assert r["stat"] < 1e-10 and r["feas"] < 1e-10 and r["dual"] == 0.0 and r["comp"] < 1e-10
    ref = so.minimize(lambda z: 0.5 * z @ H @ z + q @ z, np.zeros(n), jac=lambda z: H @ z + q, method="SLSQP",
                      constraints=[dict(type="ineq", fun=lambda z: -(G @ z + g), jac=lambda z: -G)],
                      options=dict(ftol=1e-14, maxiter=500))
    assert np.abs(ref.x - x).max() < 1e-5


def test_qp_degenerate_and_infeasible():
    H, q = np.eye(2), np.array([-1.0, -1.0])
    # duplicated constraint: x0 <= 0.5 twice -> multipliers not unique, the primal solution is
    G = np.array([[1.0, 0], [1.0, 0], [0, 1.0]])
    g = np.array([-0.5, -0.5, -0.25])
    x, lam = solve_qp_gi(H, q, G, g)
    assert np.allclose(x, [0.5, 0.25]) and np.all(lam >= 0) and np.isclose(lam[:2].sum(), 0.5)
    with pytest.raises(QPFailure):
        solve_qp_gi(H, q, np.array([[1.0, 0], [-1.0, 0]]), np.array([1.0, 1.0]))     # x <= -1 and x >= 1


def test_nearest_pd():
    rng = np.random.default_rng(0)
    A = rng.normal(size=(9, 9))
    P = nearest_pd(A)
    w = np.linalg.eigvalsh(P)
    assert np.allclose(P, P.T) and w.min() > 0
    sym = (A + A.T) / 2
    s, U = np.linalg.eigh(sym)
    assert np.allclose(P, sym + (U[:, s < 0] * (1e-10 - s[s < 0])) @ U[:, s < 0].T)
    Apd = A @ A.T + np.eye(9)
    assert np.allclose(nearest_pd(Apd), Apd)


def test_merit_directional_derivative(chicane_small):
    """The stationarity part of f_dphi (DGSQP.py:964-965) is the derivative of 1/2|q+G'l|^2 + 1/2(l.g)^2
    along (du, dl): check against central differences of the true merit."""
    og, _, _ = chicane_small
    rng = np.random.default_rng(1)
    x0 = np.array([0.5, 0.3, 2.5, 0.0, 0.5, 0.3, 1.3, -0.3, 2.2, 0.0, 1.3, -0.3])
    u = rng.normal(size=og.n) * 0.1
    l = np.abs(rng.normal(size=og.m)) * 0.1
    up = np.zeros(og.n_u)
    sol = OracleDGSQP(og, merit_function="stat")
    Q, q, G, g, _ = og.evaluate(u, l, x0, up, True)
    du, dl = rng.normal(size=og.n) * 0.1, rng.normal(size=og.m) * 0.1
    s = np.minimum(0, g)
    d0 = sol._dphi(du, l, dl, s, Q, q, G, g, 0.0)

    def phi(a):
        q2, G2, g2, _ = og.evaluate(u + a * du, l + a * dl, x0, up, False)
        return sol._phi(l + a * dl, s, q2, G2, g2, 0.0)
    h = 1e-6
    fd = (phi(h) - phi(-h)) / (2 * h)
    assert abs(fd - d0) < 1e-5 * max(1.0, abs(d0))


def test_mu_threshold_rule(chicane_small):
    og, _, _ = chicane_small
    sol = OracleDGSQP(og)
    n, m = og.n, og.m
    z, zm = np.zeros(n), np.zeros(m)
    g = -np.ones(m)
    g[3] = 1e-16                       # rounding noise at an active linear row: treated as feasible
    assert sol._get_mu(z, zm, zm, np.minimum(0, g), np.eye(n), np.ones(n), np.zeros((m, n)), g) == 0.0
    assert OracleDGSQP(og, mu_vio_thresh=0.0)._get_mu(np.ones(n), zm, zm, np.minimum(0, g), np.eye(n), np.ones(n),
                                                      np.zeros((m, n)), g) > 1e15
    g[3] = 1e-3
    assert sol._get_mu(np.ones(n), zm, zm, np.minimum(0, g), np.eye(n), np.ones(n), np.zeros((m, n)), g) > 0.0


def test_solve_regression_golden(chicane_full):
    """Full v1 solves reproduce the committed golden values (tests/golden/make_golden.py)."""
    og, _, _ = chicane_full
    data = np.load(GOLDEN / "chicane_N25_seed0.npz")
    meta = json.loads((GOLDEN / "chicane_N25_seed0.json").read_text())
    sol = OracleDGSQP(og)
    for i in meta["regression_instances"]:
        r = sol.solve(data["x0"][i], data["u_ws"][i])
        assert r["msg"] == meta["msg"][i] and r["num_iters"] == meta["num_iters"][i]
        if r["msg"] == "conv_abs_tol":
            assert np.abs(r["u"] - data["u"][i]).max() < 1e-7


# ------------------------------------------------------------------ v2 step policy (DGSQP_v2.py)
def test_v2_merit_directional_derivative(chicane_small):
    """v2 'stat_l1' (DGSQP_v2.py:1143-1161): dphi's stationarity part is the derivative of 1/2|q+G'l|^2 along
    (du, dl), with the row-stacked game Hessian Q as the Jacobian of stat w.r.t. u."""
    from oracle.dgsqp_v2 import OracleDGSQPV2
    og, _, _ = chicane_small
    rng = np.random.default_rng(2)
    x0 = np.array([0.5, 0.3, 2.5, 0.0, 0.5, 0.3, 1.3, -0.3, 2.2, 0.0, 1.3, -0.3])
    u = rng.normal(size=og.n) * 0.1
    l = np.abs(rng.normal(size=og.m)) * 0.1
    up = np.zeros(og.n_u)
    Q, q, G, g, _ = og.evaluate(u, l, x0, up, True)
    du, dl = rng.normal(size=og.n) * 0.1, rng.normal(size=og.m) * 0.1
    sol = OracleDGSQPV2(og)
    d0 = sol._dstat2(du, l, dl, Q, q, G)

    def phi(a, s=sol):
        q2, G2, _, _ = og.evaluate(u + a * du, l + a * dl, x0, up, False)
        return s._phi2(l + a * dl, np.zeros(og.m), q2, G2, 0.0)
    h = 1e-6
    assert abs((phi(h) - phi(-h)) / (2 * h) - d0) < 1e-5 * max(1.0, abs(d0))
    # 'sum_obj_l1' (:1149-1151): the smooth part is the sum of the agents' costs, its derivative grad_u(sum J)'du
    so = OracleDGSQPV2(og, merit_function="sum_obj_l1")
    og.evaluate(u, l, x0, up, True)
    d1 = so._dstat2(du, l, dl, Q, q, G)
    assert abs((phi(h, so) - phi(-h, so)) / (2 * h) - d1) < 1e-6 * max(1.0, abs(d1))
    x = og.rollout(u, x0)
    assert np.isclose(phi(0.0, so), np.sum(og.costs(x, u, up)), rtol=1e-14)


def test_v2_solve_properties_and_golden():
    """v2 solves: KKT tolerances at 'conv_abs_tol', max_it counted in m-steps, and the committed golden values."""
    from oracle.dgsqp_v2 import OracleDGSQPV2
    og = RacingGame(chicane_track(), M=2, N=15)
    data = np.load(GOLDEN / "chicane_v2_N15_seed0.npz")
    meta = json.loads((GOLDEN / "chicane_v2_N15_seed0.json").read_text())
    sol = OracleDGSQPV2(og, **meta["solver_kw"])
    for i in meta["regression_instances"]:
        r = sol.solve(data["x0"][i], data["u_ws"][i])
        assert r["msg"] == meta["msg"][i] and r["num_iters"] == meta["num_iters"][i]
        assert np.abs(r["u"] - data["u"][i]).max() < 1e-9 and np.abs(r["l"] - data["l"][i]).max() < 1e-8
        assert r["cond"]["p_feas"] < 1e-4 and r["cond"]["comp"] < 1e-4 and r["cond"]["stat"] < 1e-4
        assert r["m_steps"] + r["d_steps"] == r["num_iters"]
    # a budget of 2 m-steps cannot converge from the PID warm start: 'max_it' after exactly 2 m-steps
    r = OracleDGSQPV2(og, **dict(meta["solver_kw"], sqp_iters=2)).solve(data["x0"][0], data["u_ws"][0])
    assert r["msg"] == "max_it" and r["m_steps"] == 2 and not r["status"]
    # Newton-like setting (tiny regularisation): a handful of relaxed steps, same equilibrium as v1
    r2 = OracleDGSQPV2(og, reg=1e-3, p_tol=1e-3, d_tol=1e-3, sqp_iters=50).solve(data["x0"][0], data["u_ws"][0])
    r1 = OracleDGSQP(og).solve(data["x0"][0], data["u_ws"][0])
    assert r2["status"] and r1["status"] and np.abs(r2["u"] - r1["u"]).max() < 1e-3
