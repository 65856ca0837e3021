import sys, os, pathlib
ROOT = pathlib.Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
os.environ["DGSQP_B200_LIB"] = str(ROOT / "scripts/tmp/libdgsqp_trace.so")
import numpy as np
import dgsqp_b200 as dg
from dgsqp_b200.montecarlo import sample_head_to_head
from hostsim_lib import HostSim
inst = int(sys.argv[1])
game, params = dg.chicane_game(), dg.chicane_params()
x0, u_ws = sample_head_to_head(game, 64, seed=0)
hs = HostSim(game, params)
h = hs.solve(x0[inst], u_ws[inst])
l0 = h["l_init"]
h2 = hs.solve(x0[inst], u_ws[inst], l0)
print("HOST", h2["status"], h2["num_iters"], h2["qp_solves"], flush=True)
solver = dg.DGSQP(game, params, print_method=None)
r = solver.solve_batch(x0[inst:inst+1], u_ws[inst:inst+1], l0[None])
print("GPU", r.status[0], r.num_iters[0], r.qp_solves[0])
