"""Determinism hunt: (1) stale-workspace dependence (same CTA solves A, B, A), (2) run-to-run differences."""
import sys, pathlib
ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import numpy as np
import dgsqp_b200 as dg
from dgsqp_b200.montecarlo import sample_head_to_head

nb = int(sys.argv[1]) if len(sys.argv) > 1 else 512
threads = int(sys.argv[2]) if len(sys.argv) > 2 else 0
game, params = dg.chicane_game(), dg.chicane_params()
x0, u_ws = sample_head_to_head(game, nb, seed=0)
solver = dg.DGSQP(game, params, print_method=None, mu_vio_thresh=1e-10)
if threads: solver.configure(0, threads)
print("plan", solver.memory_plan())
fields = ("u", "l", "x", "cost", "cond", "num_iters", "status", "qp_solves")
def diff(a, b):
    return {f: int((getattr(a, f) != getattr(b, f)).sum()) for f in fields if not np.array_equal(getattr(a, f), getattr(b, f), equal_nan=True)}
# (1) one CTA, sequential single-instance calls
for (ia, ib) in ((0, 1), (2, 5), (7, 3)):
    a1 = solver.solve_batch(x0[ia:ia+1], u_ws[ia:ia+1])
    b1 = solver.solve_batch(x0[ib:ib+1], u_ws[ib:ib+1])
    a2 = solver.solve_batch(x0[ia:ia+1], u_ws[ia:ia+1])
    a3 = solver.solve_batch(x0[ia:ia+1], u_ws[ia:ia+1])
    print(f"stale test A={ia} B={ib}: A-after-B vs A-first {diff(a1, a2)}; A twice in a row {diff(a2, a3)}", flush=True)
# (2) whole batch repeated
r0 = solver.solve_batch(x0, u_ws)
for rep in range(3):
    r = solver.solve_batch(x0, u_ws)
    bad = np.where((r.u != r0.u).any(axis=1) | (r.l != r0.l).any(axis=1))[0]
    print(f"rep {rep}: {len(bad)} / {nb} instances differ; ids {bad[:10]}; max|du| {np.abs(r.u - r0.u).max():.3e} max|dl| {np.abs(r.l - r0.l).max():.3e}; "
          f"status of differing {r0.status[bad][:10]} iters {r0.num_iters[bad][:10]}", flush=True)
# (3) bit patterns and non-finite outputs
bits = lambda a: np.ascontiguousarray(a).view(np.int64)
r1 = solver.solve_batch(x0, u_ws)
print("bitwise equal u/l/x:", np.array_equal(bits(r1.u), bits(r0.u)), np.array_equal(bits(r1.l), bits(r0.l)), np.array_equal(bits(r1.x), bits(r0.x)))
nf = np.where(~np.isfinite(r0.u).all(axis=1) | ~np.isfinite(r0.l).all(axis=1))[0]
print("instances with non-finite u or l:", len(nf), nf[:20], "status", r0.status[nf][:20], "iters", r0.num_iters[nf][:20])
print("status hist", np.bincount(r0.status, minlength=6))
