"""Pins the oracle's model layer: track known answers, derivatives against torch.autograd (float64),
layout sizes against SURVEY section 8."""
import numpy as np
import pytest

from oracle.track import chicane_track, curve_track
from oracle.racing_game import RacingGame


def test_chicane_key_points_known_answer():
    # closed form of get_track_key_pts (radius_arclength_track.py:361-408) for the BASELINE chicane
    kp = chicane_track().key_pts
    want = np.array([[0, 0, 0, 0, 0, 0], [1, 0, 0, 1, 1, 0],
                     [4.601265, -1.491693, -np.pi / 4, 5, 4, -0.19635],
                     [5.308372, -2.1988, -np.pi / 4, 6, 1, 0],
                     [8.909637, -3.690493, 0, 10, 4, 0.19635],
                     [13.909637, -3.690493, 0, 15, 5, 0]])
    assert np.allclose(kp, want, atol=1e-6)


def test_track_lookups():
    t = chicane_track()
    s = np.array([0.0, 0.5, 1.0, 3.0, 5.0, 5.5, 6.0, 9.99, 10.0, 14.9, 15.0, 16.5, -0.5])
    kap = t.curvature(s)
    k = np.pi / 16
    assert np.allclose(kap, [0, 0, -k, -k, 0, 0, k, k, 0, 0, 0, -k, 0])
    psi, dpsi = t.tangent(s)
    assert np.allclose(dpsi, kap)                      # slope of pw_lin == curvature of the active segment
    assert np.isclose(psi[3], -k * 2.0) and np.isclose(psi[4], -np.pi / 4) and np.isclose(psi[8], 0.0, atol=1e-15)
    c = curve_track(curve_angle=np.pi / 2)
    assert np.isclose(c.track_length, 14.0) and np.isclose(c.tangent(np.array([9.0]))[0][0], np.pi / 2)


def test_local_to_global_matches_product_track():
    import dgsqp_b200 as dg
    to, tp = chicane_track(), dg.chicane_game().track
    rng = np.random.default_rng(0)
    s, ey, ep = rng.uniform(0, 14.9, 50), rng.uniform(-1, 1, 50), rng.uniform(-0.3, 0.3, 50)
    x, y, psi = tp.local_to_global((s, ey, ep))
    for i in range(50):
        xo, yo, po = to.local_to_global((s[i], ey[i], ep[i]))
        assert abs(xo - x[i]) < 1e-12 and abs(yo - y[i]) < 1e-12 and abs(po - psi[i]) < 1e-12
    assert np.allclose(tp.key_pts, to.key_pts)


@pytest.mark.parametrize("M,N,n,m,nc", [(2, 25, 100, 525, (16, 21, 5)), (3, 25, 150, 825, (24, 33, 9)),
                                         (4, 25, 200, 1150, (32, 46, 14))])
def test_layout_sizes(M, N, n, m, nc):
    import dgsqp_b200 as dg
    g = RacingGame(curve_track(curve_angle=np.pi / 2), M=M, N=N)
    assert (g.n, g.m) == (n, m)
    assert (g.n_c[0], g.n_c[1], g.n_c[-1]) == nc
    p = dg.agents_game(M=M, N=N)
    assert (p.n, p.m, p.n_c) == (n, m, g.n_c)


@pytest.mark.parametrize("M", [2, 3])
def test_evaluate_against_autograd(M):
    """Q, q, G, g of the oracle (DP Hessian, DGSQP.py:679-934) == direct differentiation of the batch
    Lagrangian (the reference's own unused cross-check f_Du_L / f_Duu_L, DGSQP.py:937-941)."""
    import torch_game
    rng = np.random.default_rng(M)
    g = RacingGame(chicane_track(), M=M, N=6, obs_r=0.3)
    x0 = np.concatenate([[0.5 + 0.4 * a, 0.3 - 0.3 * a, 2.5 - 0.2 * a, 0.05, 0.5 + 0.4 * a, 0.3 - 0.3 * a]
                         for a in range(M)])
    u = rng.normal(size=g.n) * 0.2
    l = np.abs(rng.normal(size=g.m))
    up = np.zeros(g.n_u)
    Q, q, G, gg, _ = g.evaluate(u, l, x0, up, True)
    Q2, q2, G2, g2 = torch_game.evaluate_autograd(g, u, l, x0, up)
    assert np.abs(Q - Q2).max() < 1e-11 * max(1.0, np.abs(Q2).max())
    assert np.abs(q - q2).max() < 1e-12 and np.abs(G - G2).max() < 1e-12 and np.abs(gg - g2).max() < 1e-13


def test_rollout_across_segment_boundary_and_wrap():
    g = RacingGame(chicane_track(), M=2, N=10)
    x0 = np.array([13.0, 0, 6.0, 0, 14.2, 0.1, 12.0, 0, 5.0, 0, 13.0, -0.2])   # s wraps past L = 15
    x = g.rollout(np.zeros(g.n), x0)
    assert np.all(np.isfinite(x)) and x[-1, 4] > 15.0
