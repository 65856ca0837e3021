"""Stale-workspace hunt: NaN-poison the CTA workspace before every instance (DGSQP_POISON=1) and compare with a plain run."""
import os, sys, pathlib
ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import numpy as np
import dgsqp_b200 as dg
from dgsqp_b200.montecarlo import sample_head_to_head
nb = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
game, params = dg.chicane_game(), dg.chicane_params()
x0, u_ws = sample_head_to_head(game, nb, seed=0)
solver = dg.DGSQP(game, params, print_method=None, mu_vio_thresh=1e-10)
bits = lambda a: np.ascontiguousarray(a).view(np.int64)
os.environ["DGSQP_POISON"] = "0"
r0 = solver.solve_batch(x0, u_ws)
os.environ["DGSQP_POISON"] = "1"
rp = solver.solve_batch(x0, u_ws)
rp2 = solver.solve_batch(x0, u_ws)
os.environ["DGSQP_POISON"] = "0"
def rep(a, b, name):
    bad = np.where((bits(a.u) != bits(b.u)).any(axis=1) | (bits(a.l) != bits(b.l)).any(axis=1) | (a.status != b.status) | (a.num_iters != b.num_iters))[0]
    print(f"{name}: {len(bad)} / {nb} instances differ: {bad[:12]}; status a {a.status[bad][:12]} b {b.status[bad][:12]}; iters a {a.num_iters[bad][:12]} b {b.num_iters[bad][:12]}")
    if len(bad): print("   max |du|", np.nanmax(np.abs(a.u[bad] - b.u[bad])), "nan in b.u:", int(np.isnan(b.u[bad]).any(axis=1).sum()), "nan in b.l:", int(np.isnan(b.l[bad]).any(axis=1).sum()))
rep(r0, rp, "plain vs poison")
rep(rp, rp2, "poison vs poison")
print("status hist plain", np.bincount(r0.status, minlength=6), "poison", np.bincount(rp.status, minlength=6))
