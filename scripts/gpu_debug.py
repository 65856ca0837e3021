"""Debug helper: solve a few chicane instances with the default plan and with shared memory disabled."""
import sys, pathlib
ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import numpy as np
import dgsqp_b200 as dg
from dgsqp_b200.montecarlo import sample_head_to_head
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
limit = int(sys.argv[2]) if len(sys.argv) > 2 else 0
game, params = dg.chicane_game(), dg.chicane_params()
x0, u_ws = sample_head_to_head(game, B, seed=0)
solver = dg.DGSQP(game, params, print_method=None, mu_vio_thresh=1e-10)
if limit:
    solver.set_smem_limit(limit)
print("plan", solver.memory_plan())
r = solver.solve_batch(x0, u_ws)
print("status", r.status, "iters", r.num_iters, "qp", r.qp_solves)
print("diag", solver.last_diag(B)[:4])
