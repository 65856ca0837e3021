"""Merge scenario (BASELINE config 5, scripts/DGSQP_merge_monte_carlo.py): oracle pins (layout, autograd, sampler),
golden fixtures against the oracle, and the CUDA kernel source compiled for the host (tests/hostsim, -DDG_GAME_MERGE)
against the oracle.  The -m gpu counterparts live in test_gpu_parity.py."""
import json
import pathlib

import numpy as np
import pytest

import dgsqp_b200 as dg
from dgsqp_b200.montecarlo import sample_merge
from hostsim_lib import HostSim
from oracle.dgsqp_v1 import OracleDGSQP
from oracle.merge_game import MergeGame, sample_merge as oracle_sample_merge, LANE

GOLDEN = pathlib.Path(__file__).parent / "golden"
MSG = {0: "conv_abs_tol", 1: "conv_rel_tol", 2: "max_it", 3: "diverged", 4: "qp_fail"}


def _rel(a, b):
    return np.abs(a - b).max() / max(1.0, np.abs(b).max())


def test_layout_sizes_survey_8():
    # merge (3 unicycles) N = 20: n_q 12, n_u 6, n = 120, n_c = 18 / 27 / 15, m = 546
    og, pg = MergeGame(), dg.merge_game()
    for g in (og, pg):
        assert (g.n_q, g.n_u, g.n, g.m) == (12, 6, 120, 546)
        assert (g.n_c[0], g.n_c[1], g.n_c[-1]) == (18, 27, 15)
    # row order inside a stage (DGSQP.py:809-820): shared, then per agent [lane x2, in-ub, in-lb, v-ub, v-lb]
    kinds = [r[1] for r in og.rows[18:18 + 27]]
    assert kinds == [0, 0, 0] + [1, 1, 2, 2, 3, 3, 4, 5] * 3


def test_lane_rows_known_answers():
    """Lane half-planes (merge.py:40-74): straight lane r <= y <= lw - r; the ramp rows switch to the straight lane's
    normals at x6 / x7 (pw_const)."""
    g = MergeGame(N=2)
    th = np.pi / 12
    x = np.zeros((3, 12))
    x[:, 0:2] = [0.3, 0.17]          # car 1 on the straight lane
    x[:, 4:6] = [1.0, 0.05]          # car 2 below the lower boundary + r
    x6x = 1.5 + 0.3 / np.tan(th)
    x[0, 8:10] = [0.5, -0.2]         # car 3 on the ramp, before the switch points
    x[1, 8:10] = [x6x + 0.02, 0.12]  # past x6 but before x7: left row already uses the straight normal
    x[2, 8:10] = [4.0, 0.15]         # past both
    val = g.constraints(x, np.zeros(g.n), np.zeros(g.n_u))
    rows = {(k, a, j): val[r] for r, (k, kind, a, j) in enumerate(g.rows) if kind == LANE}
    assert np.isclose(rows[(0, 0, 0)], 0.17 - 0.2) and np.isclose(rows[(0, 0, 1)], 0.1 - 0.17)
    assert np.isclose(rows[(0, 1, 1)], 0.1 - 0.05) and rows[(0, 1, 1)] > 0
    nm = np.array([-np.sin(th), np.cos(th)])
    x6, x7 = np.array([x6x, 0.3]), np.array([1.5 + 0.3 / np.sin(th), 0.0])
    p = x[0, 8:10]
    assert np.isclose(rows[(0, 2, 0)], nm @ (p - x6) + 0.1) and np.isclose(rows[(0, 2, 1)], -nm @ (p - x7) + 0.1)
    assert np.isclose(rows[(1, 2, 0)], 0.12 - 0.3 + 0.1)                       # straight normal (0, 1) about x6
    assert np.isclose(rows[(1, 2, 1)], -nm @ (x[1, 8:10] - x7) + 0.1)         # still the ramp normal
    assert np.isclose(rows[(2, 2, 0)], 0.15 - 0.2) and np.isclose(rows[(2, 2, 1)], 0.1 - 0.15)


def test_rk3_unicycle_known_answer():
    """rk3 (dynamics_models.py:202-212) of the unicycle with constant inputs: v and psi are exact, the position is
    Simpson's rule of v(t)[cos, sin](psi(t)) over the step."""
    g = MergeGame()
    q, u, h = np.array([0.3, -0.2, 0.7, 0.4]), np.array([1.2, -2.0]), g.dt
    qn = g.fd_agent(q, u)
    v = lambda t: q[2] + t * u[0] / g.mass
    psi = lambda t: q[3] + t * u[1]
    simpson = lambda f: h / 6 * (f(0.0) + 4 * f(h / 2) + f(h))
    assert np.isclose(qn[2], v(h)) and np.isclose(qn[3], psi(h))
    assert np.isclose(qn[0], q[0] + simpson(lambda t: v(t) * np.cos(psi(t))), atol=1e-15)
    assert np.isclose(qn[1], q[1] + simpson(lambda t: v(t) * np.sin(psi(t))), atol=1e-15)


def test_evaluate_against_autograd():
    """Q, q, G, g of the oracle (DP Hessian) == direct differentiation of the batch Lagrangian."""
    import torch_merge
    g = MergeGame(N=5)
    rng = np.random.default_rng(2)
    x0 = oracle_sample_merge(3, seed=1, game=g)[2]
    x0[8:10] = [2.55, 0.1]       # car 3 near the switch points so that both lane normals occur along the horizon
    x0[10] = 1.5
    u = rng.normal(size=g.n) * 0.5
    l = np.abs(rng.normal(size=g.m))
    Q, q, G, gg, x = g.evaluate(u, l, x0, np.zeros(g.n_u), True)
    assert x[:, 8].min() < g.lanes[2][0][0] < x[:, 8].max()
    Q2, q2, G2, g2 = torch_merge.evaluate_autograd(g, u, l, x0)
    assert np.abs(Q - Q2).max() < 1e-11 * max(1.0, np.abs(Q2).max())
    assert np.abs(q - q2).max() < 1e-12 * max(1.0, np.abs(q2).max())
    assert np.abs(G - G2).max() < 1e-12 and np.abs(gg - g2).max() < 1e-13


def test_host_sampler_matches_oracle_sampler():
    """dgsqp_b200.montecarlo.sample_merge (vectorised) draws the script's stream in the script's order."""
    a, u = sample_merge(dg.merge_game(), 300, seed=1)
    b = oracle_sample_merge(300, seed=1)
    assert np.array_equal(a, b) and not u.any()
    # the script's pre-check (car 3 rolled out from the zero state) rejects some trials
    r = np.random.default_rng(1).random((300, 12))
    assert not np.allclose(a[:, 0], 0.5 * r[:, 0] - 0.25)


def test_golden_matches_oracle():
    data = np.load(GOLDEN / "merge_N20_seed1.npz")
    meta = json.loads((GOLDEN / "merge_N20_seed1.json").read_text())
    og = MergeGame(N=20)
    assert np.array_equal(data["x0"], oracle_sample_merge(32, seed=1, game=og))
    sol = OracleDGSQP(og, reg=0.0)
    for i in meta["regression_instances"] + [5]:
        r = sol.solve(data["x0"][i], data["u_ws"][i])
        assert r["msg"] == meta["msg"][i] and r["num_iters"] == meta["num_iters"][i]
        assert _rel(r["u"], data["u"][i]) < 1e-9 and _rel(r["l"], data["l"][i]) < 1e-8


@pytest.mark.parametrize("N,smem", [(20, None), (7, "0")])
def test_kernel_source_evaluate_and_G_products(N, smem, monkeypatch):
    if smem is not None:
        monkeypatch.setenv("DG_HOSTSIM_SMEM_DOUBLES", smem)      # everything in the global workspace
    og = MergeGame(N=N)
    hs = HostSim(dg.merge_game(N=N), dg.merge_params(N))
    assert (hs.nq, hs.nu, hs.n, hs.m) == (og.n_q, og.n_u, og.n, og.m)
    rng = np.random.default_rng(N)
    x0 = oracle_sample_merge(2, seed=3, game=og)[1]
    x0[8:10] = [2.5, 0.1]
    x0[10] = 1.2
    u = rng.normal(size=og.n) * 0.5
    l = np.abs(rng.normal(size=og.m)) * (rng.random(og.m) < 0.4)
    Q, q, G, g, x = og.evaluate(u, l, x0, np.zeros(og.n_u), True)
    Q2, q2, gtl2, g2, x2 = hs.evaluate(x0, u, l)
    assert np.abs(x - x2).max() < 1e-13 and np.abs(g - g2).max() < 1e-13 and np.abs(q - q2).max() < 1e-12
    assert np.abs(G.T @ l - gtl2).max() < 1e-12 and np.abs(Q - Q2).max() < 1e-12 * max(1.0, np.abs(Q).max())
    assert np.abs(G - hs.G_dense()).max() < 1e-13
    v, w = rng.normal(size=og.n), rng.normal(size=og.m)
    assert np.abs(G @ v - hs.G_times(v)).max() < 1e-12 and np.abs(G.T @ w - hs.GT_times(w)).max() < 1e-12


def test_kernel_source_solves_vs_golden():
    """Full solves of the kernel source (own LSQR dual initialisation) against the oracle's golden results: identical
    status, iteration and QP counts; trajectories x and costs within 1e-6.  The script sets reg = 0, so the QP Hessian
    nearestPD(Q) keeps eigenvalues at the 1e-10 floor (condition ~1e11): inputs along those directions agree to
    cond * eps ~ 1e-5 and the multipliers to 1e-4, which is what is asserted (measured: u <= 2.5e-6, l <= 1.4e-5)."""
    data = np.load(GOLDEN / "merge_N20_seed1.npz")
    meta = json.loads((GOLDEN / "merge_N20_seed1.json").read_text())
    hs = HostSim(dg.merge_game(), dg.merge_params())
    same = 0
    for i in range(12):
        r = hs.solve(data["x0"][i], data["u_ws"][i])
        assert MSG[r["status"]] == meta["msg"][i] and _rel(r["l_init"], data["l_init"][i]) < 1e-9
        # with reg = 0 the iteration path of a rare instance is sensitive to summation order (the KKT test at 1e-3 is
        # met one or two iterations earlier or later); measured: 159 of 160 instances identical (golden 32 + 128 fresh)
        if r["num_iters"] != meta["num_iters"][i]:
            assert abs(r["num_iters"] - meta["num_iters"][i]) <= 2 and _rel(r["x"].ravel(), data["x"][i]) < 1e-4
            continue
        same += 1
        assert r["qp_solves"] == meta["qp_solves"][i]
        assert _rel(r["u"], data["u"][i]) < 1e-5 and _rel(r["x"].ravel(), data["x"][i]) < 1e-6
        assert _rel(r["l"], data["l"][i]) < 1e-4 and _rel(r["cost"], data["cost"][i]) < 1e-6
    assert same >= 11


def test_merge_v2_policy_kernel_source_vs_oracle():
    """DGSQPV2Params step policy on the merge game: kernel source == oracle restatement of DGSQP_v2.solve."""
    from oracle.dgsqp_v2 import OracleDGSQPV2
    N = 10
    og = MergeGame(N=N)
    kw = dict(reg=1e-1, reg_decay=0.8, nms_frequency=3, sqp_iters=40, p_tol=1e-4, d_tol=1e-4)
    params = dg.DGSQPV2Params(N=N, **kw)
    hs = HostSim(dg.merge_game(N=N), params)
    sol = OracleDGSQPV2(og, **kw)
    X0 = oracle_sample_merge(3, seed=1, game=og)
    for i in range(3):
        ref = sol.solve(X0[i], np.zeros(og.n))
        r = hs.solve(X0[i], np.zeros(og.n))
        assert MSG[r["status"]] == ref["msg"] and r["num_iters"] == ref["num_iters"]
        if ref["status"]:
            assert _rel(r["u"], ref["u"]) < 1e-6


def test_host_class_surface_merge():
    """The product-side game record: state2q order, struct marshalling, error behaviour (no GPU needed)."""
    g = dg.merge_game()
    st = [dg.VehicleState(x=dg.Position(x=0.1 * a, y=0.15), e=dg.OrientationEuler(psi=0.01 * a),
                          v=dg.BodyLinearVelocity(v_long=0.3)) for a in range(3)]
    q = g.state2q(st)
    assert q.shape == (12,) and np.allclose(q[4:8], [0.1, 0.15, 0.3, 0.01])
    s = g.to_struct()
    assert s.M == 3 and s.N == 20 and s.lane[2][0].brk == pytest.approx(1.5 + 0.3 / np.tan(np.pi / 12))
    assert s.lane[0][0].brk == float("inf") and s.goal[1][0] == 4.5 and s.term_scale == 10.0
    with pytest.raises(ValueError):
        dg.MergeGame(M=3, obs_r=[0.1, 0.1])


@pytest.mark.parametrize("M,idx", [(2, [0, 2]), (4, [0, 1, 2, 2])])
def test_merge_game_other_agent_counts(M, idx):
    """The merge record is not tied to three cars: two (one straight-lane car + the ramp car) and four agents, kernel
    source == oracle for the evaluation and for a full solve (same status; the 4-car start is infeasible: 'qp_fail')."""
    from dgsqp_b200.games import merge_lanes as product_lanes
    from oracle.merge_game import merge_lanes as oracle_lanes
    N = 8
    goals = [(4.0, .15, .3, 0), (4.5, .15, .3, 0), (4.25, .15, .3, 0), (3.5, .15, .3, 0)][:M]
    pl, ol = product_lanes(), oracle_lanes()
    g = dg.MergeGame(M=M, N=N, goals=goals, obs_r=[0.1] * M, lanes=[pl[i] for i in idx])
    og = MergeGame(M=M, N=N, goals=goals, obs_r=[0.1] * M, lanes=[ol[i] for i in idx])
    assert (g.n, g.m, g.n_c) == (og.n, og.m, og.n_c)
    hs = HostSim(g, dg.DGSQPParams(N=N, reg=1e-3))
    x3 = oracle_sample_merge(1, seed=1)[0].reshape(3, 4)
    x0 = np.concatenate([x3[i] + (0.0 if k < 3 else np.array([-0.6, 0, 0, 0])) for k, i in enumerate(idx)])
    rng = np.random.default_rng(0)
    u, l = rng.normal(size=og.n) * 0.3, np.abs(rng.normal(size=og.m)) * 0.2
    Q, q, G, gg, x = og.evaluate(u, l, x0, np.zeros(og.n_u))
    Qh, qh, gtl, gh, xh = hs.evaluate(x0, u, l)
    assert np.abs(Q - Qh).max() < 1e-12 and np.abs(q - qh).max() < 1e-12 and np.abs(gg - gh).max() < 1e-13
    assert np.abs(G - hs.G_dense()).max() < 1e-13 and np.abs(G.T @ l - gtl).max() < 1e-12
    r, h = OracleDGSQP(og, reg=1e-3).solve(x0, np.zeros(og.n)), hs.solve(x0, np.zeros(og.n))
    assert MSG[h["status"]] == r["msg"] and h["num_iters"] == r["num_iters"]
    if r["status"]:
        assert _rel(h["u"], r["u"]) < 1e-8
