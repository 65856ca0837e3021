#!/usr/bin/env python3
"""Headline benchmark: converged 2-agent chicane game solves / s (BASELINE.json configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B] [--workload chicane|merge]

A step = one solve_batch over B = 10 000 synthetic chicane instances per GPU (randomised initial conditions,
PID warm start; dgsqp_b200.montecarlo).  `value` is measured with the inputs resident in HBM (CUDA events on
the launching stream), `e2e` through the host-buffer C-ABI call with the H2D / D2H copies inside the timed
region.  N > 1: one process per GPU under torchrun, instances sharded with no collective on the solve path
(weak scaling: B per GPU), statistics gathered at the end, time = max over ranks.
`--impl reference` times the CPU oracle (NumPy restatement of the reference solver; the reference itself
needs CasADi + OSQP which are not installable offline) on all host cores.
`--workload merge` runs the same measurement on the merge scenario (BASELINE.json configs[4]: three unicycles,
N = 20, seed-1 sampler of scripts/DGSQP_merge_monte_carlo.py); the default is the headline chicane workload.
"""
import argparse
import json
import os
import pathlib
import sys
import threading
import time

ROOT = pathlib.Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "converged_game_solves_per_sec"
UNIT = "solves/s"
WORKLOAD = "chicane_2agent_N25_mc"


# ----------------------------------------------------------------------------- CPU baseline (oracle port)
def _oracle_worker(args):
    x0, u_ws = args
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    import numpy as np  # noqa: F401
    from oracle.dgsqp_v1 import OracleDGSQP
    global _ORACLE
    try:
        _ORACLE
    except NameError:
        if len(x0) == 12 and len(u_ws) == 120:          # merge scenario (3 unicycles, N = 20)
            from oracle.merge_game import MergeGame
            _ORACLE = OracleDGSQP(MergeGame(N=20), reg=0.0)
        else:
            from oracle.racing_game import RacingGame
            from oracle.track import chicane_track
            _ORACLE = OracleDGSQP(RacingGame(chicane_track(), M=2, N=25))
    r = _ORACLE.solve(x0, u_ws)
    return bool(r["status"]), int(r["num_iters"]), r["msg"], r["u"]


def cpu_oracle_throughput(x0, u_ws, cores):
    """Solve the given sample with one oracle instance per core; returns (converged/s, iters/s, seconds)."""
    import multiprocessing as mp
    for k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[k] = "1"
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        pool.map(_oracle_worker, [(x0[i], u_ws[i]) for i in range(min(cores, len(x0)))])      # warm the workers
        t0 = time.perf_counter()
        out = pool.map(_oracle_worker, [(x0[i], u_ws[i]) for i in range(len(x0))], chunksize=1)
        dt = time.perf_counter() - t0
    conv = sum(o[0] for o in out)
    iters = sum(o[1] for o in out)
    global _LAST_ORACLE_RESULTS
    _LAST_ORACLE_RESULTS = out          # (status, iters, msg, u) per instance: parity of the GPU path on the same sample
    return conv / dt, iters / dt, dt, conv


def _native_worker(args):
    x0, u_ws, merge = args
    sys.path.insert(0, str(ROOT / "tests"))
    import dgsqp_b200 as dg
    from hostsim_lib import HostSim
    global _NATIVE
    try:
        _NATIVE
    except NameError:
        _NATIVE = HostSim(dg.merge_game(), dg.merge_params()) if merge else HostSim(dg.chicane_game(), dg.chicane_params())
    r = _NATIVE.solve(x0, u_ws)
    return int(r["status"]) <= 1, int(r["num_iters"])


def cpu_native_throughput(x0, u_ws, cores, merge):
    """The SAME solver source compiled for the host (tests/hostsim: single-thread C++ build of csrc/*.cuh), one
    instance per core.  Not the reference and not the product: an honest yardstick for what a CPU does with this
    algorithm once Python/NumPy overhead is gone."""
    import multiprocessing as mp
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        pool.map(_native_worker, [(x0[i], u_ws[i], merge) for i in range(min(cores, len(x0)))])
        t0 = time.perf_counter()
        out = pool.map(_native_worker, [(x0[i], u_ws[i], merge) for i in range(len(x0))], chunksize=2)
        dt = time.perf_counter() - t0
    return sum(o[0] for o in out) / dt, len(out) / dt, sum(o[1] for o in out) / dt, dt


# ----------------------------------------------------------------------------- clocks sampler
class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag = index, [], set(), False
        self.max_mhz = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                 nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap"}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        import statistics
        return dict(sm_mhz=(statistics.median(self.samples) if self.samples else None), sm_max_mhz=self.max_mhz,
                    reasons=sorted(self.reasons), samples=len(self.samples))


# ----------------------------------------------------------------------------- algorithmic work model
def algorithmic_flops(game, diag, num_qp, lsqr_iters=20):
    """FP64 flops of the path by the counts the kernel reports ('Measurement'; conventions of
    SURVEY 8(d): multiply-add = 2 flops, dense formulas without credit for agent block sparsity)."""
    import numpy as np
    n, m, N, nq, nu, M = game.n, game.m, game.N, game.n_q, game.n_u, game.M
    F_jac = N * N * nq * nq * nu + m * nq * n
    F_hess = (M + 1) * N * N * nu * nq * (nq + nu)
    F_pd = 4.0 / 3.0 * n ** 3                   # Householder tridiagonalisation (negative eigenpairs are O(n^2) each)
    F_chol = 2.0 / 3.0 * n ** 3                 # Cholesky + triangular inverse
    F_gi = 5.0 * n * n + 2.0 * 3 * M * N * N    # J'n, z, Householder update, slack evaluation
    F_lsqr = lsqr_iters * (8.0 * 3 * M * N * N + 8.0 * lsqr_iters * m)
    full, grad, gi = (diag[:, k].astype(np.float64) for k in range(3))
    return float((full * (F_jac + F_hess) + grad * F_jac + num_qp * (F_pd + F_chol) + gi * F_gi + F_lsqr).sum())


def algorithmic_bytes(game, B):
    """HBM bytes the path has to move: inputs + outputs per instance (~8.5 KB at chicane size)."""
    per = 8 * (game.n_q + game.n + (game.n + game.m + (game.N + 1) * game.n_q + game.M + 3)) + 3 * 4
    return per * B


# ----------------------------------------------------------------------------- main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=10000, help="instances per GPU per step")
    ap.add_argument("--cpu-sample", type=int, default=0, help="instances for the CPU baseline (0 = 1 per core)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="chicane", choices=["chicane", "merge"])
    args = ap.parse_args()
    # stdout carries exactly one JSON line: NCCL's own prints (version banner, NCCL_DEBUG output) go to stderr
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    cores = os.cpu_count() or 1

    import numpy as np
    import dgsqp_b200 as dg
    from dgsqp_b200.montecarlo import sample_head_to_head, sample_merge

    if args.workload == "merge":
        game, params = dg.merge_game(), dg.merge_params()
        sample = lambda g, B, seed: sample_merge(g, B, seed=1 + seed)     # the script's seed is 1
        config = dict(workload="merge_3agent_N20_mc", instances_per_gpu=args.batch, agents=3, horizon=20, n=game.n,
                      m=game.m, solver="DGSQP v1 (DGSQPParams: reg=0, nonmono_ls, 50 SQP iters, tol 1e-3)",
                      l2_policy="256 MiB buffer written between timed steps (L2 flush)", sampler_seed=1)
    else:
        game, params = dg.chicane_game(), dg.chicane_params()
        sample = sample_head_to_head
        config = dict(workload=WORKLOAD, instances_per_gpu=args.batch, agents=2, horizon=25, n=game.n, m=game.m,
                      solver="DGSQP v1 (DGSQPParams: reg=1e-3, nonmono_ls, 50 SQP iters, tol 1e-3)",
                      l2_policy="256 MiB buffer written between timed steps (L2 flush)", sampler_seed=0)

    # ------------------------------------------------------------------ reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return
        n_s = args.cpu_sample or cores
        x0, u_ws = sample(game, n_s * (args.steps + args.warmup), 0)
        vals, its, secs = [], [], []
        for s in range(args.warmup + args.steps):
            sl = slice(s * n_s, (s + 1) * n_s)
            v, ips, dt, _ = cpu_oracle_throughput(x0[sl], u_ws[sl], cores)
            if s >= args.warmup:
                vals.append(v); its.append(ips); secs.append(dt)
        value = float(np.mean(vals))
        line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                    ms_per_step=1e3 * float(np.mean(secs)), higher_is_better=True, scaling="weak", vs_baseline=None,
                    dtype="f64", data="synthetic", config=dict(config, instances_per_step=n_s), impl="reference",
                    sqp_iters_per_sec=float(np.mean(its)),
                    cpu_baseline=dict(value=value, unit=UNIT, cores=cores, kind="port",
                                      sample=f"{n_s} {args.workload} instances per step, one oracle process per core"),
                    e2e=dict(value=value, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
        print(json.dumps(line))
        return

    # ------------------------------------------------------------------ our arm (GPU)
    import torch
    from dgsqp_b200 import _abi
    from dgsqp_b200.sharding import shard_stats, gather_stats
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: dgsqp_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    B = args.batch
    x0, u_ws = sample(game, B, rank)                             # each rank owns its shard of the global batch
    solver = dg.DGSQP(game, params, print_method=None, device=local_rank)
    lib = _abi.load()
    x0_d, u_d = torch.from_numpy(x0).to(dev), torch.from_numpy(u_ws).to(dev)
    out = solver.alloc_outputs(B, dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream(dev)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def device_step():
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        res = solver._solve_batch_device(x0_d, u_d, None, stream.cuda_stream, out, sync=False)
        e1.record(stream)
        return res, e0, e1

    for _ in range(args.warmup):
        device_step()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = lib.dgsqp_kernel_launches()
    barrier()
    evs = []
    for _ in range(args.steps):
        res, e0, e1 = device_step()
        evs.append((e0, e1))
    barrier()
    launches = lib.dgsqp_kernel_launches() - launches0
    kernel_ms = [e0.elapsed_time(e1) for e0, e1 in evs]
    sampler.stop_flag = True
    sampler.join(timeout=2)
    status = res.status.cpu().numpy(); iters = res.num_iters.cpu().numpy(); qps = res.qp_solves.cpu().numpy()
    cond = res.cond.cpu().numpy()
    diag = solver.last_diag(B)
    t_dev = sum(kernel_ms) / 1e3

    # end-to-end through the host-buffer C-ABI call (pinned host memory, copies inside the timed region)
    pin = lambda a: torch.from_numpy(a).pin_memory().numpy()
    x0_p, u_p = pin(x0), pin(u_ws)
    solver.solve_batch(x0_p, u_p)
    barrier()
    e2e_secs = []
    for _ in range(args.steps):
        flush.fill_(1)
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        r_h = solver.solve_batch(x0_p, u_p)
        e2e_secs.append(time.perf_counter() - t0)
    barrier()
    t_e2e = sum(e2e_secs)
    assert np.array_equal(r_h.status, status), "host path and device path disagree"
    h2d = x0_p.nbytes + u_p.nbytes
    d2h = sum(a.nbytes for a in (r_h.u, r_h.l, r_h.x, r_h.cost, r_h.cond, r_h.num_iters, r_h.status, r_h.qp_solves))

    # max over ranks, totals over ranks
    conv_local = int((status <= 1).sum())
    t_dev_max, t_e2e_max, conv_tot, iters_tot = t_dev, t_e2e, conv_local, int(iters.sum())
    if dist is not None:
        tt = torch.tensor([t_dev, t_e2e], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_dev_max, t_e2e_max = float(tt[0]), float(tt[1])
        cc = torch.tensor([conv_local, int(iters.sum())], dtype=torch.float64, device=dev)
        dist.all_reduce(cc, op=dist.ReduceOp.SUM)
        conv_tot, iters_tot = int(cc[0]), int(cc[1])
    stats = gather_stats(shard_stats(status, iters, qps, cond))

    if rank == 0:
        K = args.steps
        value = conv_tot * K / t_dev_max
        peak_tf = C_double_peak(lib, local_rank)
        fl = algorithmic_flops(game, diag, qps.astype(np.float64))
        ach_tf = fl * K / t_dev / 1e12
        peaks = {}
        try:
            peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
        except Exception:
            pass
        # DRAM traffic of the dominant kernel from the committed `ncu --set full` capture (per instance, scaled to the
        # launch): profiles/ncu_traffic.json, written from the raw page of the same kernel on one wave of instances
        traffic, traffic_src = None, None
        try:
            tj = json.loads((ROOT / "profiles" / "ncu_traffic.json").read_text()).get(config["workload"])
            if tj:
                traffic, traffic_src = tj["dram_bytes_per_instance"] * B, tj["source"]
        except Exception:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        hbm_ach = algorithmic_bytes(game, B) * K / t_dev / 1e9
        line = dict(
            metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=K, warmup=args.warmup,
            ms_per_step=1e3 * t_dev_max / K, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f64",
            data="synthetic", config=config, sqp_iters_per_sec=iters_tot * K / t_dev_max,
            solves_per_sec_all=stats["count"] * K / t_dev_max,
            e2e=dict(value=conv_tot * K / t_e2e_max, unit=UNIT, h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h),
            gpu_launches=int(launches),
            roofline=dict(bound="fp64", achieved=ach_tf, peak=peak_tf, unit="TFLOP/s",
                          frac=(ach_tf / peak_tf if peak_tf else None), traffic=traffic, traffic_source=traffic_src,
                          note="peak = FP64 FMA probe kernel measured in this run (MEASURED_PEAKS.json has no FP64 "
                               "figure); achieved = algorithmic flops from per-instance work counters / "
                               "CUDA-event time of dgsqp_solve_kernel"),
            roofline_hbm=dict(bound="hbm", achieved=hbm_ach, peak=hbm_peak, unit="GB/s", frac=hbm_ach / hbm_peak,
                              traffic=None, peak_source="measured" if "hbm_gbs" in peaks else "fallback"),
            clocks=sampler.summary(),
            stats={k: stats[k] for k in ("count", "converged", "conv_abs_tol", "conv_rel_tol", "max_it", "diverged",
                                         "qp_fail", "mean_iters", "std_iters", "sum_qp")},
            work=dict(full_evals=float(diag[:, 0].mean()), grad_evals=float(diag[:, 1].mean()),
                      qp_active_set_iters=float(diag[:, 2].mean()), algorithmic_mflop_per_instance=fl / B / 1e6))
        if not args.no_cpu_baseline:
            n_s = args.cpu_sample or cores
            v, ips, dt, conv = cpu_oracle_throughput(x0[:n_s], u_ws[:n_s], cores)
            line["cpu_baseline"] = dict(value=v, unit=UNIT, cores=cores, kind="port",
                                        sample=f"first {n_s} instances of the same batch, one oracle process per core, "
                                               f"{dt:.1f} s wall", sqp_iters_per_sec=ips)
            # parity of the GPU results with the oracle on that sample (status, iteration count; u on converged ones)
            msgs = _abi.STATUS_MSG
            u_gpu = res.u[:n_s].cpu().numpy()
            same, worst = 0, 0.0
            for i, (st_o, it_o, msg_o, u_o) in enumerate(_LAST_ORACLE_RESULTS):
                if msgs[int(status[i])] == msg_o and int(iters[i]) == it_o:
                    same += 1
                    if st_o:
                        worst = max(worst, float(np.abs(u_gpu[i] - u_o).max() / max(1.0, np.abs(u_o).max())))
            try:
                n_n = 16 * cores
                cv, av, iv, ndt = cpu_native_throughput(x0[:n_n], u_ws[:n_n], cores, args.workload == "merge")
                line["cpu_native"] = dict(value=cv, unit=UNIT, solves_per_sec_all=av, sqp_iters_per_sec=iv, cores=cores,
                                          kind="kernel source compiled for the host (tests/hostsim, single-thread C++ "
                                               "build of the same solver), one instance per core",
                                          sample=f"first {n_n} instances of the same batch, {ndt:.1f} s wall")
            except Exception as e:                      # the yardstick is optional (needs g++ artefacts of the tests)
                line["cpu_native"] = dict(unavailable=str(e)[:200])
            line["parity"] = dict(sample=n_s, identical_status_and_iters=same, max_rel_err_u_converged=worst,
                                  against="oracle port (own LSQR dual initialisation on both sides)")
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def C_double_peak(lib, device):
    import ctypes
    v = ctypes.c_double(0.0)
    rc = lib.dgsqp_measure_fp64_peak(int(device), ctypes.byref(v))
    return float(v.value) if rc == 0 else None


if __name__ == "__main__":
    main()
