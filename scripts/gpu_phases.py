"""Per-phase cycle profile of the solve kernel (dgsqp_last_phase_cycles) for a chicane batch."""
import sys, time, pathlib, json
ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import numpy as np
import torch
import dgsqp_b200 as dg
from dgsqp_b200 import _abi
from dgsqp_b200.montecarlo import sample_head_to_head

B = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
threads = int(sys.argv[2]) if len(sys.argv) > 2 else 0
ctas = int(sys.argv[3]) if len(sys.argv) > 3 else 0
import os
if os.environ.get("DG_WORKLOAD") == "merge":
    from dgsqp_b200.montecarlo import sample_merge
    game, params = dg.merge_game(), dg.merge_params()
    x0, u_ws = sample_merge(game, B, seed=1)
elif os.environ.get("DG_WORKLOAD", "").startswith("agents"):
    from dgsqp_b200.montecarlo import sample_agents
    M = int(os.environ["DG_WORKLOAD"][6:])
    game, params = dg.agents_game(M=M, N=25), dg.agents_params(25)
    x0, u_ws = sample_agents(game, B, seed=0)
elif os.environ.get("DG_WORKLOAD", "").startswith("curve"):
    th = float(os.environ["DG_WORKLOAD"][5:] or 45)
    game, params = dg.curve_game(th, 25), dg.curve_params(25)
    x0, u_ws = sample_head_to_head(game, B, seed=1)
else:
    game, params = dg.chicane_game(), dg.chicane_params()
    x0, u_ws = sample_head_to_head(game, B, seed=0)
solver = dg.DGSQP(game, params, print_method=None, mu_vio_thresh=1e-10)
if threads or ctas:
    solver.configure(ctas, threads)
dev = torch.device("cuda:0")
x0d, ud = torch.from_numpy(x0).to(dev), torch.from_numpy(u_ws).to(dev)
for rep in range(2):
    torch.cuda.synchronize(); t = time.time()
    r = solver.solve_batch(x0d, ud); torch.cuda.synchronize(); el = time.time() - t
st = r.status.cpu().numpy(); it = r.num_iters.cpu().numpy(); qp = r.qp_solves.cpu().numpy()
print(os.environ.get("DG_WORKLOAD", "chicane"), f"B {B} threads {threads} ctas/SM {ctas}: {el:.3f} s -> {B/el:.1f} solves/s, converged {int((st<=1).sum())}, iters/s {it.sum()/el:.0f}")
print("status hist", np.bincount(st, minlength=5), "mean iters", it.mean(), "mean qp", qp.mean())
d = solver.last_diag(B)
print("diag mean [full evals, grad evals, GI iters, max nneg, indef QPs, sum nneg, sum active, ls trials]:", d.mean(axis=0), "max", d.max(axis=0))
ph = solver.last_phase_cycles(B).astype(np.float64)
tot = ph.sum()
print("phase shares (of CTA cycles) and mean kcycles per instance:")
for k, name in enumerate(_abi.PHASES):
    if ph[:, k].sum() > 0 or k < 13:
        print(f"  {name:12s} {100*ph[:,k].sum()/tot:6.2f} %   {ph[:,k].mean()/1e3:10.1f} kcyc   per QP {ph[:,k].sum()/max(qp.sum(),1)/1e3:8.1f} kcyc")
print(f"  total mean {ph.sum(axis=1).mean()/1e6:.2f} Mcyc per instance; per full eval hess {ph[:,2].sum()/max(d[:,0].sum(),1)/1e3:.1f} kcyc; "
      f"per QP tridiag {ph[:,3].sum()/max(qp.sum(),1)/1e3:.1f} eig {ph[:,4].sum()/max(qp.sum(),1)/1e3:.1f} chol {ph[:,5].sum()/max(qp.sum(),1)/1e3:.1f} "
      f"trinv {ph[:,6].sum()/max(qp.sum(),1)/1e3:.1f} GI {ph[:,7].sum()/max(qp.sum(),1)/1e3:.1f} (per GI it {ph[:,7].sum()/max(d[:,2].sum(),1)/1e3:.2f}) kcyc; "
      f"per grad eval {(ph[:,9]+ph[:,10]).sum()/max(d[:,1].sum(),1)/1e3:.1f} kcyc")
