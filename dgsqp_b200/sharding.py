"""Instance sharding across the GPUs of one node and the final statistics gather.

Every Monte-Carlo instance is an independent game (the reference drivers are plain sequential loops,
e.g. ``scripts/DGSQP_merge_monte_carlo.py:428``), so the batch is split contiguously by instance index,
one process per GPU, with NO collective on the solve path.  Only the per-shard statistics -- the table
``scripts/process_data_curve.py:98-110`` / ``process_data_merge.py:58-67`` prints (converged / failed /
max-it counts, mean and std of SQP iterations, QP solves) -- are gathered at the end with one
``all_gather`` of 16 doubles.
"""
import numpy as np

STAT_KEYS = ["count", "conv_abs_tol", "conv_rel_tol", "max_it", "diverged", "qp_fail", "time_limit",
             "sum_iters", "sum_iters_sq", "sum_qp", "sum_conv_iters", "sum_conv_iters_sq", "sum_conv_qp",
             "max_p_feas_conv", "max_stat_conv", "reserved"]


def shard_bounds(total: int, world_size: int, rank: int):
    """Contiguous shard [lo, hi) of `total` instances for `rank`; sizes differ by at most one."""
    base, rem = divmod(total, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_stats(status, num_iters, qp_solves, cond=None) -> np.ndarray:
    """Additive statistics vector (see STAT_KEYS) of one shard; max-type entries are kept separately."""
    status, num_iters, qp_solves = (np.asarray(a) for a in (status, num_iters, qp_solves))
    v = np.zeros(len(STAT_KEYS))
    v[0] = status.size
    for code in range(6):
        v[1 + code] = int((status == code).sum())
    it = num_iters.astype(np.float64)
    v[7], v[8], v[9] = it.sum(), (it ** 2).sum(), qp_solves.sum()
    conv = status <= 1
    v[10], v[11], v[12] = it[conv].sum(), (it[conv] ** 2).sum(), qp_solves[conv].sum()
    if cond is not None and conv.any():
        cond = np.asarray(cond)
        v[13], v[14] = cond[conv, 0].max(), cond[conv, 2].max()
    return v


def combine_stats(vectors) -> dict:
    """Merge per-shard vectors into the reference's summary table."""
    vs = np.asarray(vectors, dtype=np.float64).reshape(-1, len(STAT_KEYS))
    tot = vs.sum(axis=0)
    tot[13], tot[14] = vs[:, 13].max(), vs[:, 14].max()
    out = {k: (float(tot[i]) if k.startswith(("sum", "max")) else int(tot[i])) for i, k in enumerate(STAT_KEYS) if k != "reserved"}
    n, nc = max(out["count"], 1), max(out["conv_abs_tol"] + out["conv_rel_tol"], 1)
    out["converged"] = out["conv_abs_tol"] + out["conv_rel_tol"]
    out["mean_iters"] = out["sum_iters"] / n
    out["std_iters"] = float(np.sqrt(max(out["sum_iters_sq"] / n - out["mean_iters"] ** 2, 0.0)))
    out["mean_conv_iters"] = out["sum_conv_iters"] / nc
    out["std_conv_iters"] = float(np.sqrt(max(out["sum_conv_iters_sq"] / nc - out["mean_conv_iters"] ** 2, 0.0)))
    return out


def gather_stats(local_vector: np.ndarray) -> dict:
    """all_gather of the per-shard statistics over the default process group (NCCL or gloo);
    single-process runs skip the collective."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return combine_stats([local_vector])
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    t = torch.as_tensor(local_vector, dtype=torch.float64, device=dev)
    out = [torch.empty_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    return combine_stats([o.cpu().numpy() for o in out])
