"""Small fixed workload for ncu: one wave of chicane instances through the device path."""
import sys, pathlib
ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import numpy as np, torch
import dgsqp_b200 as dg
from dgsqp_b200.montecarlo import sample_head_to_head
B = int(sys.argv[1]) if len(sys.argv) > 1 else 296
import os
if os.environ.get("DG_WORKLOAD") == "merge":
    from dgsqp_b200.montecarlo import sample_merge
    game, params = dg.merge_game(), dg.merge_params()
    x0, u_ws = sample_merge(game, B, seed=1)
else:
    game, params = dg.chicane_game(), dg.chicane_params()
    x0, u_ws = sample_head_to_head(game, B, seed=0)
solver = dg.DGSQP(game, params, print_method=None, mu_vio_thresh=1e-10)
dev = torch.device("cuda:0")
r = solver.solve_batch(torch.from_numpy(x0).to(dev), torch.from_numpy(u_ws).to(dev))
torch.cuda.synchronize()
print("done", int((r.status <= 1).sum()), "converged of", B)
