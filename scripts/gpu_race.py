"""Small run for compute-sanitizer racecheck / memcheck: selected instances of the seed-0 chicane batch."""
import sys, pathlib
ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import numpy as np
import dgsqp_b200 as dg
from dgsqp_b200.montecarlo import sample_head_to_head
ids = [int(v) for v in sys.argv[1].split(",")] if len(sys.argv) > 1 else [0, 1]
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 50
game, params = dg.chicane_game(), dg.chicane_params()
params.sqp_iters = iters
x0, u_ws = sample_head_to_head(game, max(ids) + 1, seed=0)
solver = dg.DGSQP(game, params, print_method=None, mu_vio_thresh=1e-10)
r = solver.solve_batch(x0[ids], u_ws[ids])
print("status", r.status, "iters", r.num_iters, "qp", r.qp_solves)
