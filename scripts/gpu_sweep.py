"""Occupancy sweep of the solve kernel: threads per CTA x CTAs per SM x shared-memory cap (dgsqp_configure /
dgsqp_set_smem_limit).  With a cap below the two n x n work matrices the planner keeps pool + sensitivities in shared
memory and the matrices in the L2-resident global workspace, which lets several CTAs (instances) share an SM.
    python scripts/gpu_sweep.py [chicane|merge] [B]"""
import sys, time, pathlib
ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import numpy as np
import torch
import dgsqp_b200 as dg
from dgsqp_b200.montecarlo import sample_head_to_head, sample_merge

wl = sys.argv[1] if len(sys.argv) > 1 else "chicane"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 2960
if wl == "merge":
    game, params = dg.merge_game(), dg.merge_params()
    x0, u_ws = sample_merge(game, B, seed=1)
else:
    game, params = dg.chicane_game(), dg.chicane_params()
    x0, u_ws = sample_head_to_head(game, B, seed=0)
dev = torch.device("cuda:0")
x0d, ud = torch.from_numpy(x0).to(dev), torch.from_numpy(u_ws).to(dev)
ref = None
configs = [(256, 1, 0), (256, 1, 110), (128, 1, 0), (128, 2, 110), (128, 2, 80), (64, 2, 110), (64, 4, 56), (64, 3, 75),
           (96, 2, 110), (192, 1, 0), (128, 1, 110), (64, 1, 0)]
for threads, ctas, cap_kb in configs:
    solver = dg.DGSQP(game, params, print_method=None, mu_vio_thresh=1e-10)
    try:
        if cap_kb:
            solver.set_smem_limit(cap_kb * 1024)
        solver.configure(ctas, threads)
        plan = solver.memory_plan()
        best = 1e9
        for rep in range(2):
            torch.cuda.synchronize(); t = time.time()
            r = solver.solve_batch(x0d, ud); torch.cuda.synchronize()
            best = min(best, time.time() - t)
        st = r.status.cpu().numpy(); it = r.num_iters.cpu().numpy()
        if ref is None:
            ref = (st.copy(), it.copy())
        same = float(np.mean((st == ref[0]) & (it == ref[1])))
        print(f"{wl} B {B} threads {threads:3d} ctas/SM {ctas} cap {cap_kb:3d} KB | smem {plan['smem_bytes']/1024:6.1f} KB mats_in_smem {int(plan['mats_in_smem'])} "
              f"sens {int(plan['sens_in_smem'])} | {best:.3f} s  {B/best:8.1f} solves/s  conv {int((st<=1).sum())}  same-as-first {same:.4f}", flush=True)
    except Exception as e:
        print(f"{wl} threads {threads} ctas {ctas} cap {cap_kb}: FAILED {e}", flush=True)
    del solver
