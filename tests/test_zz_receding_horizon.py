"""Receding-horizon use of the v2 solver class on the GPU (first step): u_prev through the C ABI
(dgsqp_solve_batch_up) and step() keeping u_prev / shifting the warm start like DGSQP_v2.py:301-320."""
import numpy as np
import pytest

import dgsqp_b200 as dg
from dgsqp_b200.montecarlo import sample_head_to_head

pytestmark = pytest.mark.gpu


def test_v2_u_prev_and_step_sequence():
    from oracle.dgsqp_v2 import OracleDGSQPV2
    from oracle.racing_game import RacingGame
    from oracle.track import chicane_track
    N = 10
    kw = dict(reg=1e-2, reg_decay=0.8, nms_frequency=2, sqp_iters=40, p_tol=1e-4, d_tol=1e-4)
    game, og = dg.chicane_game(N=N), RacingGame(chicane_track(), M=2, N=N)
    solver, sol = dg.DGSQP(game, dg.DGSQPV2Params(N=N, **kw), print_method=None, mu_vio_thresh=1e-10), OracleDGSQPV2(og, **kw)
    x0, u_ws = sample_head_to_head(game, 6, seed=4)
    rng = np.random.default_rng(0)
    up = np.column_stack([rng.uniform(-1.5, 1.5, 6), rng.uniform(-0.3, 0.3, 6), rng.uniform(-1.5, 1.5, 6), rng.uniform(-0.3, 0.3, 6)])
    res = solver.solve_batch(x0, u_ws, u_prev=up)
    base = solver.solve_batch(x0, u_ws)
    assert np.abs(res.u - base.u).max() > 1e-3
    for i in range(6):
        r = sol.solve(x0[i], u_ws[i], u_prev=up[i])
        assert res.msg[i] == r["msg"] and int(res.num_iters[i]) == r["num_iters"]
        if r["status"]:
            assert np.abs(res.u[i] - r["u"]).max() < 1e-6 * max(1.0, np.abs(r["u"]).max())
    import torch
    dev = torch.device("cuda:0")
    rd = solver.solve_batch(torch.from_numpy(x0).to(dev), torch.from_numpy(u_ws).to(dev), u_prev=torch.from_numpy(up).to(dev))
    assert np.array_equal(rd.u.cpu().numpy(), res.u)
    # step(): the second solve sees u_prev = first applied input (DGSQP_v2.py:311) and the shifted warm start (:313-316)
    states = []
    for a in range(2):
        s = dg.VehicleState(t=0.0)
        s.x.x, s.x.y, s.v.v_long, s.p.e_psi, s.p.s, s.p.x_tran = x0[0, 6 * a:6 * a + 6]
        states.append(s)
    solver.set_warm_start(solver.agent_to_stage_major(u_ws[:1])[0])
    info1 = solver.step(states)
    u1 = solver.u_pred.copy()
    assert np.array_equal(solver.u_prev, u1[0])
    info2 = solver.step(states)
    r1 = sol.solve(x0[0], u_ws[0])
    assert info1["msg"] == r1["msg"] and info1["num_iters"] == r1["num_iters"]
    if info1["msg"] not in ("diverged", "qp_fail"):
        shifted = solver.stage_to_agent_major(np.vstack((u1[1:], u1[-1]))[None])[0]
        r2 = sol.solve(x0[0], shifted, u_prev=u1[0])
        assert info2["msg"] == r2["msg"] and info2["num_iters"] == r2["num_iters"]


def test_step_batch_closed_loop_matches_single_steps():
    """step_batch: B races advanced in lock-step.  Two closed-loop steps of 5 instances -- the states moved
    by the bicycle model of the host sampler between the steps -- give, per instance, exactly what the single-instance
    step() sequence of the reference-shaped class gives (same applied inputs, status, iterations)."""
    N = 10
    kw = dict(reg=1e-2, reg_decay=0.8, nms_frequency=2, sqp_iters=40, p_tol=1e-4, d_tol=1e-4)
    game = dg.chicane_game(N=N)
    x0, u_ws = sample_head_to_head(game, 5, seed=9)
    batch = dg.DGSQP(game, dg.DGSQPV2Params(N=N, **kw), print_method=None, mu_vio_thresh=1e-10)
    u0a, ra = batch.step_batch(x0, u_ws=u_ws)
    x1 = ra.x.reshape(5, N + 1, game.n_q)[:, 1]                    # the solver's own roll-out: state after applying u0
    u0b, rb = batch.step_batch(x1)
    assert u0a.shape == (5, game.n_u) and np.array_equal(u0a, batch.agent_to_stage_major(ra.u)[:, 0])
    for i in range(5):
        single = dg.DGSQP(game, dg.DGSQPV2Params(N=N, **kw), print_method=None, mu_vio_thresh=1e-10)
        single.set_warm_start(single.agent_to_stage_major(u_ws[i:i + 1])[0])
        for step, (xs, u0, res) in enumerate(((x0, u0a, ra), (x1, u0b, rb))):
            states = []
            for a in range(2):
                s = dg.VehicleState(t=0.0)
                s.x.x, s.x.y, s.v.v_long, s.p.e_psi, s.p.s, s.p.x_tran = xs[i, 6 * a:6 * a + 6]
                states.append(s)
            info = single.step(states)
            assert info["msg"] == res.msg[i] and info["num_iters"] == int(res.num_iters[i]), (i, step)
            assert np.array_equal(single.u_pred[0], u0[i]), (i, step)
