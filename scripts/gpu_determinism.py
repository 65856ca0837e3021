"""Run the same batch with different CTA widths and repeatedly; report agreement with the host build."""
import sys, time, pathlib
ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np
import dgsqp_b200 as dg
from dgsqp_b200.montecarlo import sample_head_to_head
from hostsim_lib import HostSim

nchk = int(sys.argv[1]) if len(sys.argv) > 1 else 64
game, params = dg.chicane_game(), dg.chicane_params()
x0, u_ws = sample_head_to_head(game, nchk, seed=0)
hs = HostSim(game, params)
H = [hs.solve(x0[i], u_ws[i]) for i in range(nchk)]
hst = np.array([h["status"] for h in H]); hit = np.array([h["num_iters"] for h in H]); hqp = np.array([h["qp_solves"] for h in H])
solver = dg.DGSQP(game, params, print_method=None, mu_vio_thresh=1e-10)
# same dual initialisation for both sides (isolates the LSQR chaos)
l0 = np.stack([h["l_init"] for h in H])
H2 = [hs.solve(x0[i], u_ws[i], l0[i]) for i in range(nchk)]
h2st = np.array([h["status"] for h in H2]); h2it = np.array([h["num_iters"] for h in H2])
print("host: lsqr vs given-l0 self check", ((h2st == hst) & (h2it == hit)).sum(), "/", nchk)
for threads in (32, 128, 256):
    solver.configure(0, threads)
    r = solver.solve_batch(x0, u_ws, l0)
    same = (r.status == h2st) & (r.num_iters == h2it)
    conv = h2st == 0
    errs = [np.abs(r.u[i] - H2[i]["u"]).max() for i in range(nchk) if same[i] and h2st[i] <= 1]
    print(f"given l0, threads {threads}: identical {same.sum()}/{nchk}; on host-conv_abs subset {same[conv].sum()}/{conv.sum()}; max|du| on matching converged {max(errs):.2e}")
    print("   mismatches:", [(int(i), int(h2st[i]), int(h2it[i]), int(r.status[i]), int(r.num_iters[i])) for i in np.where(~same)[0]])
prev = None
for threads in (32, 64, 128, 128, 256):
    solver.configure(0, threads)
    t = time.time(); r = solver.solve_batch(x0, u_ws); el = time.time() - t
    same = (r.status == hst) & (r.num_iters == hit)
    sameqp = same & (r.qp_solves == hqp)
    msg = f"threads {threads:4d}: {el:.2f}s  identical(status,iters) {same.sum()}/{nchk}  (+qp) {sameqp.sum()}"
    if prev is not None:
        msg += f"   vs previous run: status/iters equal {((r.status == prev.status) & (r.num_iters == prev.num_iters)).sum()}/{nchk}, max|du| {np.abs(r.u - prev.u).max():.2e}"
    print(msg, flush=True)
    prev = r
