"""Ad-hoc GPU check: CUDA path vs single-thread host build of the same source, plus timing."""
import sys, time, pathlib
ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np
import torch
import dgsqp_b200 as dg
from dgsqp_b200.montecarlo import sample_head_to_head
from hostsim_lib import HostSim

nchk = int(sys.argv[1]) if len(sys.argv) > 1 else 32
nbig = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
game, params = dg.chicane_game(), dg.chicane_params()
x0, u_ws = sample_head_to_head(game, max(nchk, nbig), seed=0)
solver = dg.DGSQP(game, params, print_method=None, mu_vio_thresh=1e-10)
t = time.time(); res = solver.solve_batch(x0[:nchk], u_ws[:nchk]); print("gpu solve", nchk, "instances:", time.time() - t, "s")
hs = HostSim(game, params)
same = 0; worst = 0.0
for i in range(nchk):
    h = hs.solve(x0[i], u_ws[i])
    ok = h["status"] == res.status[i] and h["num_iters"] == res.num_iters[i]
    same += ok
    err = np.abs(h["u"] - res.u[i]).max()
    if h["status"] <= 1 and ok: worst = max(worst, err)
    if not ok or (h["status"] <= 1 and err > 1e-6):
        print("  inst", i, "host", h["status"], h["num_iters"], h["qp_solves"], "gpu", res.status[i], res.num_iters[i], res.qp_solves[i], "du", err)
print(f"identical (status, iters): {same}/{nchk}; worst |du| on converged: {worst:.2e}")
print("status hist", np.bincount(res.status, minlength=5), "mean iters", res.num_iters.mean())
dev = torch.device("cuda:0")
x0d, ud = torch.from_numpy(x0[:nbig]).to(dev), torch.from_numpy(u_ws[:nbig]).to(dev)
for rep in range(2):
    torch.cuda.synchronize(); t = time.time()
    r = solver.solve_batch(x0d, ud); torch.cuda.synchronize(); el = time.time() - t
    st = r.status.cpu().numpy(); it = r.num_iters.cpu().numpy()
    print(f"device batch {nbig}: {el:.3f} s -> {nbig/el:.1f} solves/s, converged {int((st<=1).sum())}, iters/s {it.sum()/el:.0f}")
d = solver.last_diag(nbig)
print("diag mean [full evals, grad evals, GI iters, max nneg]:", d.mean(axis=0), "max", d.max(axis=0))
