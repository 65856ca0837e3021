#!/usr/bin/env python3
"""Generates the golden fixtures from the CPU oracle (the reference itself cannot run offline:
CasADi / OSQP are not importable -- see oracle/__init__.py).  Instances come from the oracle's
faithful sampler (SciPy RK45 PID roll-outs, seed recorded).

    python tests/golden/make_golden.py [chicane] [curve] [agents3] [chicane_v2] [merge]
"""
import json
import pathlib
import sys
import time

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle.track import chicane_track, curve_track          # noqa: E402
from oracle.racing_game import RacingGame                    # noqa: E402
from oracle.sampler import sample_head_to_head, sample_agents  # noqa: E402
from oracle.dgsqp_v1 import OracleDGSQP                      # noqa: E402
from oracle.dgsqp_v2 import OracleDGSQPV2                    # noqa: E402
from oracle.merge_game import MergeGame, sample_merge        # noqa: E402

OUT = pathlib.Path(__file__).resolve().parent


def make(name, game, sampler, solver_kw, count, seed, regression, cls=OracleDGSQP):
    rng = np.random.default_rng(seed)
    sol = cls(game, **solver_kw)
    keys = ["x0", "u_ws", "l_init", "u", "l", "x", "cost", "cond"]
    arr = {k: [] for k in keys}
    meta = dict(name=name, seed=seed, count=count, msg=[], num_iters=[], qp_solves=[], solver_kw=solver_kw,
                regression_instances=regression, generator="tests/golden/make_golden.py (CPU oracle)")
    t0 = time.time()
    for i in range(count):
        x0, u_ws = sampler(game, rng)
        r = sol.solve(x0, u_ws)
        vals = [x0, u_ws, r["init"]["l"], r["u"], r["l"], r["x"].ravel(), r["cost"],
                [r["cond"]["p_feas"], r["cond"]["comp"], r["cond"]["stat"]]]
        for k, v in zip(keys, vals):
            arr[k].append(np.asarray(v, dtype=np.float64))
        meta["msg"].append(r["msg"])
        meta["num_iters"].append(int(r["num_iters"]))
        meta["qp_solves"].append(int(r["qp_solves"]))
        print(f"{name} {i}: {r['msg']} {r['num_iters']} ({time.time() - t0:.0f}s)", flush=True)
    np.savez_compressed(OUT / f"{name}.npz", **{k: np.stack(v) for k, v in arr.items()})
    (OUT / f"{name}.json").write_text(json.dumps(meta, indent=1))


if __name__ == "__main__":
    which = sys.argv[1:] or ["chicane", "curve", "agents3", "chicane_v2", "merge"]
    if "merge" in which:
        # scripts/DGSQP_merge_monte_carlo.py: seed 1, zero warm start, reg = 0; the sampler is sequential in the script's order
        g = MergeGame(N=20)
        X0 = iter(sample_merge(32, seed=1, game=g))
        make("merge_N20_seed1", g, lambda game, rng: (next(X0), np.zeros(game.n)), dict(reg=0.0), 32, 1, [0])
    if "chicane" in which:
        make("chicane_N25_seed0", RacingGame(chicane_track(), M=2, N=25), sample_head_to_head, dict(reg=1e-3), 48, 0,
             [3, 4, 7])
    if "curve" in which:
        g = RacingGame(curve_track(curve_angle=np.pi / 4), M=2, N=15, rate_ub=(10.0, 4.5), rate_lb=(-10.0, -4.5),
                       obs_r=0.2)
        make("curve45_N15_seed1", g, sample_head_to_head, dict(reg=0.0), 24, 1, [0])
    if "agents3" in which:
        g = RacingGame(curve_track(curve_angle=np.pi / 2), M=3, N=15, obs_r=0.4)
        make("agents3_N15_seed0", g, sample_agents, dict(reg=1e-3), 16, 0, [0])
    if "chicane_v2" in which:
        # v2 step policy (DGSQPV2Params); a faster regularisation decay than the class default keeps the file small
        make("chicane_v2_N15_seed0", RacingGame(chicane_track(), M=2, N=15), sample_head_to_head,
             dict(reg=1e2, reg_decay=0.9, nms_frequency=3, sqp_iters=60, p_tol=1e-4, d_tol=1e-4), 16, 0, [0],
             cls=OracleDGSQPV2)
