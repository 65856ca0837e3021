#!/usr/bin/env python3
"""Round-2 golden fixtures (CPU oracle, same conventions as make_golden.py) for the configurations VERDICT r1 asked GPU
parity tests for: four agents at N = 25 (n = 200, m = 1150), the 75 and 90 degree curves at N = 25, and 128 more merge
instances (accepted samples 32..159 of the script's seeded sampler).  Instances are solved by a process pool.

    python tests/golden/make_golden_r2.py [agents4] [curve75] [curve90] [merge_b] [--procs 4]
"""
import json
import multiprocessing as mp
import os
import pathlib
import sys
import time

for _k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
    os.environ[_k] = "1"
import numpy as np  # noqa: E402

ROOT = pathlib.Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
OUT = pathlib.Path(__file__).resolve().parent
_SOL = {}


def _game(kind):
    from oracle.track import curve_track
    from oracle.racing_game import RacingGame
    from oracle.merge_game import MergeGame
    if kind == "agents4":
        return RacingGame(curve_track(curve_angle=np.pi / 2), M=4, N=25, obs_r=0.4), dict(reg=1e-3)
    if kind in ("curve75", "curve90"):
        ang = (75.0 if kind == "curve75" else 90.0) * np.pi / 180
        return RacingGame(curve_track(curve_angle=ang), M=2, N=25, rate_ub=(10.0, 4.5), rate_lb=(-10.0, -4.5), obs_r=0.2), dict(reg=0.0)
    return MergeGame(N=20), dict(reg=0.0)


def _solve(args):
    kind, i, x0, u_ws = args
    from oracle.dgsqp_v1 import OracleDGSQP
    if kind not in _SOL:
        g, kw = _game(kind)
        _SOL[kind] = OracleDGSQP(g, **kw)
    t0 = time.time()
    r = _SOL[kind].solve(x0, u_ws)
    print(f"{kind} {i}: {r['msg']} {r['num_iters']} ({time.time() - t0:.0f}s)", flush=True)
    return dict(i=i, l_init=r["init"]["l"], u=r["u"], l=r["l"], x=r["x"].ravel(), cost=np.asarray(r["cost"], float),
                cond=[r["cond"]["p_feas"], r["cond"]["comp"], r["cond"]["stat"]], msg=r["msg"], num_iters=int(r["num_iters"]),
                qp_solves=int(r["qp_solves"]))


def make(pool, name, kind, count, seed, skip=0):
    from oracle.sampler import sample_head_to_head, sample_agents
    from oracle.merge_game import sample_merge
    g, kw = _game(kind)
    if kind == "merge_b":
        X0 = sample_merge(skip + count, seed=seed, game=g)[skip:]
        inst = [(X0[i], np.zeros(g.n)) for i in range(count)]
    else:
        rng = np.random.default_rng(seed)
        sampler = sample_agents if kind == "agents4" else sample_head_to_head
        inst = [sampler(g, rng) for _ in range(count)]
    rs = pool.map(_solve, [(kind, i, x0, u) for i, (x0, u) in enumerate(inst)], chunksize=1)
    arr = dict(x0=np.stack([x for x, _ in inst]), u_ws=np.stack([u for _, u in inst]))
    for k in ("l_init", "u", "l", "x", "cost", "cond"):
        arr[k] = np.stack([np.asarray(r[k], dtype=np.float64) for r in rs])
    meta = dict(name=name, seed=seed, count=count, skip=skip, msg=[r["msg"] for r in rs], num_iters=[r["num_iters"] for r in rs],
                qp_solves=[r["qp_solves"] for r in rs], solver_kw=kw, generator="tests/golden/make_golden_r2.py (CPU oracle)")
    np.savez_compressed(OUT / f"{name}.npz", **arr)
    (OUT / f"{name}.json").write_text(json.dumps(meta, indent=1))


if __name__ == "__main__":
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    procs = int(sys.argv[sys.argv.index("--procs") + 1]) if "--procs" in sys.argv else 4
    which = args or ["curve75", "curve90", "merge_b", "agents4"]
    with mp.get_context("fork").Pool(procs) as pool:
        if "curve75" in which:
            make(pool, "curve75_N25_seed1", "curve75", 12, 1)
        if "curve90" in which:
            make(pool, "curve90_N25_seed1", "curve90", 12, 1)
        if "merge_b" in which:
            make(pool, "merge_N20_seed1_b", "merge_b", 128, 1, skip=32)
        if "agents4" in which:
            make(pool, "agents4_N25_seed0", "agents4", 12, 0)
