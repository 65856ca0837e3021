"""Batched Monte-Carlo driver (dgsqp_b200/drivers.py) with a CPU stand-in for the solver: the single-thread host build of
the kernel source behind the solve_batch surface (test harness only)."""
import numpy as np

import dgsqp_b200 as dg
from dgsqp_b200.drivers import run_monte_carlo, summary_table
from dgsqp_b200.solver import BatchResult
from hostsim_lib import HostSim


class _HostSolver:
    def __init__(self, game, params):
        self.game, self.hs = game, HostSim(game, params)

    def solve_batch(self, x0, u_ws):
        rs = [self.hs.solve(x0[i], u_ws[i]) for i in range(len(x0))]
        st = lambda k, dt=np.float64: np.array([r[k] for r in rs], dtype=dt)
        return BatchResult(st("u"), st("l"), np.stack([r["x"].ravel() for r in rs]), st("cost"), st("cond"),
                           st("num_iters", np.int32), st("status", np.int32), st("qp_solves", np.int32), 0.0)


def test_monte_carlo_driver_records_and_table():
    game, params = dg.merge_game(N=10), dg.merge_params(10)
    records, stats = run_monte_carlo(_HostSolver(game, params), 10, seed=1, chunk=4, keep_trajectories=True)
    assert len(records) == 10 and stats["count"] == 10
    info = records[0]["solve_info"]
    assert set(info) >= {"time", "num_iters", "status", "msg", "qp_solves", "cost", "cond"} and set(info["cond"]) == {"p_feas", "comp", "stat"}
    assert records[3]["q"].shape == (11, 12) and records[3]["init"]["x0"].shape == (12,)
    lines = []
    tab = summary_table(records, "merge N: 10", lines.append)
    assert tab["converged"] == stats["converged"] == sum(r["solve_info"]["status"] for r in records)
    assert any(l.startswith("Converged") for l in lines) and any(l.startswith("Avg solves") for l in lines)
    assert abs(tab["avg_iters"] - stats["mean_conv_iters"]) < 1e-12
