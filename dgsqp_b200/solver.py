"""Host-side DG-SQP solver class over the CUDA C-ABI.

Mirrors the surface the reference's Monte-Carlo drivers use on ``DGSQP.solvers.DGSQP.DGSQP``
(``DGSQP/solvers/DGSQP.py``): ``set_warm_start`` (:271-281), ``solve`` (:302-507), ``step``
(:283-297), ``get_prediction`` (:299-300) and the attributes ``q_pred``, ``u_pred``, ``l_pred``,
``n_c``, ``N``, ``M`` -- plus ``solve_batch``, which is the reason this package exists: thousands of
independent instances in one launch.  All arithmetic happens in ``libdgsqp_b200.so``; this module
only marshals buffers.  There is no CPU fallback.
"""
import array
import copy
import ctypes as C
import time
from typing import List, Optional

import numpy as np

from . import _abi
from .games import RacingGame, MergeGame, params_to_struct, params_v2_to_struct, NUA
from .solver_types import DGSQPParams, DGSQPV2Params
from .types import VehicleState, VehiclePrediction


class BatchResult:
    """Arrays returned by :meth:`DGSQP.solve_batch` (NumPy for host calls, torch CUDA tensors for
    device calls).  ``msg`` decodes ``status`` into the reference's strings."""

    def __init__(self, u, l, x, cost, cond, num_iters, status, qp_solves, elapsed):
        self.u, self.l, self.x, self.cost, self.cond = u, l, x, cost, cond
        self.num_iters, self.status, self.qp_solves, self.elapsed = num_iters, status, qp_solves, elapsed

    @property
    def msg(self):
        st = self.status.cpu().numpy() if hasattr(self.status, "cpu") else self.status
        return [_abi.STATUS_MSG[int(s)] for s in st]

    @property
    def converged(self):
        return self.status <= 1


MU_VIO_REFERENCE = 0.0        # `thresh` of DGSQP._get_mu (DGSQP.py:560): the literal reference rule, the class default
MU_VIO_DETERMINISTIC = 1e-10  # the penalty-weight switch no longer reacts to +-1 ulp at active linear rows


class DGSQP:
    """``DGSQP(game, params)`` or, like the reference (DGSQP.py:25-33), ``DGSQP(joint_dynamics, costs, agent_constraints,
    shared_constraints, bounds, params)``.  ``game`` is one of the records of :mod:`dgsqp_b200.games` (the literals the
    reference scripts feed CasADi); the six-argument form is routed through
    :func:`dgsqp_b200.frontend.game_from_reference_args`, which identifies that record from plain-Python cost / constraint
    callables and raises ``NotImplementedError`` for anything outside the supported game family.

    ``mu_vio_thresh``: 0 (default) is the reference's rule; ``MU_VIO_DETERMINISTIC`` is the setting every parity fixture
    was generated with.  ``qp_warm_start``: each QP of an instance starts from the active set of its
    previous QP (same solution, fewer active-set iterations); False = cold start like the reference (DGSQP.py:240-241)."""

    def __init__(self, game, *args, **kw):
        from .dynamics import CasadiDecoupledMultiAgentDynamicsModel
        if isinstance(game, CasadiDecoupledMultiAgentDynamicsModel):
            # the reference's own form (DGSQP.py:25-33):
            #   DGSQP(joint_dynamics, costs, agent_constraints, shared_constraints, bounds, params=..., print_method=...)
            from .frontend import game_from_reference_args
            names = ["costs", "agent_constraints", "shared_constraints", "bounds", "params", "print_method"]
            ref = dict(zip(names, args))
            for k in names:
                if k in kw:
                    ref[k] = kw.pop(k)
            missing = [k for k in names[:4] if k not in ref]
            if missing:
                raise TypeError("DGSQP(joint_dynamics, costs, agent_constraints, shared_constraints, bounds, params): missing "
                                + ", ".join(missing))
            for k in ("xy_plot", "use_mx"):          # accepted and ignored like a head-less run of the reference
                kw.pop(k, None)
            params = ref.get("params") or DGSQPParams()
            record = game_from_reference_args(game, ref["costs"], ref["agent_constraints"], ref["shared_constraints"],
                                              ref["bounds"], params)
            self._init(record, params, ref.get("print_method", print), **kw)
        else:
            self._init(game, *args, **kw)

    def _init(self, game: RacingGame, params: DGSQPParams = None, print_method=print, device: int = 0,
              mu_vio_thresh: float = MU_VIO_REFERENCE, qp_warm_start: bool = True):
        if params is None:
            params = DGSQPParams()
        self.v2 = isinstance(params, DGSQPV2Params)      # step policy of DGSQP_v2.py instead of DGSQP.py
        if params.N != game.N:
            raise ValueError("params.N = %i but the game was built for N = %i" % (params.N, game.N))
        if params.qp_solver != "osqp" or params.qp_interface != "casadi":
            raise ValueError(f"Unsupported QP interface {params.qp_interface}/{params.qp_solver}")
        if params.hessian_approximation != "none":
            raise ValueError(f"Hessian approximation method {params.hessian_approximation} not implmented")
        self.game, self.params = game, params
        self.print_method = (lambda s: None) if print_method is None else print_method
        self.M, self.N = game.M, game.N
        self.n_u, self.n_q = game.n_u, game.n_q
        self.n_c = game.n_c
        self.merge = isinstance(game, MergeGame)         # merge scenario (unicycles) instead of a racing game (bicycles)
        self.nqa = game.n_q // game.M
        self.num_qa_d = [self.nqa] * self.M
        self.num_ua_d = [NUA] * self.M
        self.num_ua_el = [self.N * NUA] * self.M
        self.solver_name = params.solver_name

        self._lib = _abi.load()
        gs = game.to_struct()
        self._h = C.c_void_p()
        self.mu_vio_thresh, self.qp_warm_start = float(mu_vio_thresh), bool(qp_warm_start)
        self.save_iter_data = bool(getattr(params, "save_iter_data", False))
        kw = dict(mu_vio_thresh=self.mu_vio_thresh, qp_warm_start=self.qp_warm_start, iter_log=self.save_iter_data)
        if self.v2:
            ps = params_v2_to_struct(params, **kw)
            create = self._lib.dgsqp_create_merge_v2 if self.merge else self._lib.dgsqp_create_v2
        else:
            ps = params_to_struct(params, **kw)
            create = self._lib.dgsqp_create_merge if self.merge else self._lib.dgsqp_create
        _abi.check(create(C.byref(gs), C.byref(ps), int(device), C.byref(self._h)))
        dims = (C.c_int32 * 4)()
        _abi.check(self._lib.dgsqp_dims(self._h, dims))
        assert (dims[0], dims[1], dims[2], dims[3]) == (game.n_q, game.n_u, game.n, game.m)
        self.device = device

        self.q_pred = np.zeros((self.N + 1, self.n_q))
        self.u_pred = np.zeros((self.N, self.n_u))
        self.l_pred = np.zeros(game.m)
        self.u_prev = np.zeros(self.n_u)
        self.u_ws = np.zeros(self.N * self.n_u)
        self.l_ws = None
        self.state_input_predictions = [VehiclePrediction() for _ in range(self.M)]
        self.initialized = True

    def __del__(self):
        h = getattr(self, "_h", None)
        if h is not None and h.value:
            self._lib.dgsqp_destroy(h)
            self._h = None

    # ------------------------------------------------------------------ reference surface
    def initialize(self):
        pass

    def set_warm_start(self, u_ws: np.ndarray, l_ws: np.ndarray = None):
        if u_ws.shape[0] != self.N or u_ws.shape[1] != self.n_u:
            raise RuntimeError('Warm start state sequence of shape (%i,%i) is incompatible with required shape (%i,%i)'
                               % (u_ws.shape[0], u_ws.shape[1], self.N, self.n_u))
        self.u_ws = self.stage_to_agent_major(np.asarray(u_ws, dtype=np.float64)[None])[0]
        self.l_ws = l_ws

    def stage_to_agent_major(self, u_ws):
        """[B, N, n_u] (agent column blocks) -> [B, n] agent-major (DGSQP.py:275-280)."""
        B = u_ws.shape[0]
        return np.ascontiguousarray(u_ws.reshape(B, self.N, self.M, NUA).transpose(0, 2, 1, 3).reshape(B, -1))

    def agent_to_stage_major(self, u):
        """[B, n] agent-major -> [B, N, n_u] (DGSQP.py:477-482)."""
        B = u.shape[0]
        return u.reshape(B, self.M, self.N, NUA).transpose(0, 2, 1, 3).reshape(B, self.N, self.n_u)

    def solve(self, states: List[VehicleState], parameters: np.ndarray = np.array([])):
        t0 = time.time()
        if not self.v2:
            self.u_prev = np.zeros(self.n_u)           # v1 zeroes u_prev in solve (DGSQP.py:305); v2 keeps it (DGSQP_v2.py:328)
        x0 = self.game.state2q(states)
        res = self.solve_batch(x0[None], self.u_ws[None], u_prev=np.asarray(self.u_prev, dtype=np.float64)[None] if self.v2 else None)
        msg = res.msg[0]
        self.q_pred = res.x[0].reshape(self.N + 1, self.n_q)
        self.u_pred = self.agent_to_stage_major(res.u)[0]
        self.l_pred = res.l[0]
        dur = time.time() - t0
        self.print_method(self.solver_name)
        self.print_method(f'Solve status: {msg}')
        self.print_method(f'Solve iters: {int(res.num_iters[0])}')
        self.print_method(f'Solve time: {dur:.2f}')
        self.print_method(str(res.cost[0]))
        cond = dict(p_feas=float(res.cond[0, 0]), comp=float(res.cond[0, 1]), stat=float(res.cond[0, 2]))
        if self.v2 and msg == "time_limit":
            msg = "time_limit_exceeded"                # the v2 class's string (DGSQP_v2.py:412)
        iter_data = self.last_iter_data(1)[0] if self.save_iter_data else []
        return dict(time=dur, num_iters=int(res.num_iters[0]), status=bool(res.status[0] <= 1), cost=res.cost[0],
                    cond=cond, iter_data=iter_data, msg=msg, init=dict(u=self.u_ws.copy(), l=None),
                    qp_solves=int(res.qp_solves[0]))

    def step(self, states: List[VehicleState], parameters: np.ndarray = np.array([])):
        info = self.solve(states, parameters)
        for a, st in enumerate(states):
            st.u.u_a, st.u.u_steer = float(self.u_pred[0, NUA * a]), float(self.u_pred[0, NUA * a + 1])
        self._fill_predictions(states[0].t)
        self.u_prev = self.u_pred[0]
        if info['msg'] not in ['diverged', 'qp_fail']:
            self.set_warm_start(np.vstack((self.u_pred[1:], self.u_pred[-1])))
        return info

    def step_batch(self, x0, u_ws=None, u_prev=None):
        """Receding-horizon step of B independent games at once: the batched form of ``step`` (DGSQP.py:283-297,
        DGSQP_v2.py:301-328) for closed-loop Monte-Carlo runs (many races advanced in lock-step).

        ``x0`` [B, n_q]; ``u_ws`` [B, n] agent-major warm start (default: the shifted solutions this solver kept from the
        previous call, zeros on the first); ``u_prev`` [B, n_u] previous inputs (v2 policy; default: what the previous
        call applied).  Returns ``(u0, result)`` -- ``u0`` [B, n_u] the first-stage inputs to apply, ``result`` the
        :class:`BatchResult` -- and keeps, per instance, ``u_prev = u0`` and the warm start shifted by one stage with the
        last stage repeated, except where the solve diverged / failed its QP (the previous warm start stays, like the
        reference)."""
        x0 = np.ascontiguousarray(x0, dtype=np.float64)
        B = x0.shape[0]
        st = getattr(self, "_batch_state", None)
        if st is None or st["u_ws"].shape[0] != B:
            st = self._batch_state = dict(u_ws=np.zeros((B, self.game.n)), u_prev=np.zeros((B, self.n_u)))
        if u_ws is not None:
            st["u_ws"] = np.ascontiguousarray(u_ws, dtype=np.float64).copy()
        if u_prev is not None:
            st["u_prev"] = np.ascontiguousarray(u_prev, dtype=np.float64).copy()
        res = self.solve_batch(x0, st["u_ws"], u_prev=st["u_prev"] if self.v2 else None)
        u_sm = self.agent_to_stage_major(res.u)                     # [B, N, n_u]
        u0 = np.ascontiguousarray(u_sm[:, 0])
        st["u_prev"] = u0.copy()
        ok = ~np.isin(res.status, (3, 4))                           # 'diverged', 'qp_fail' keep the old warm start
        shifted = self.stage_to_agent_major(np.concatenate((u_sm[:, 1:], u_sm[:, -1:]), axis=1))
        st["u_ws"][ok] = shifted[ok]
        return u0, res

    def get_prediction(self) -> List[VehiclePrediction]:
        return self.state_input_predictions

    def _fill_predictions(self, t):
        if self.merge:
            # CasadiKinematicUnicycle.qu2prediction (dynamics_models.py:347-359)
            for a, pred in enumerate(self.state_input_predictions):
                q = self.q_pred[:, 4 * a:4 * (a + 1)]
                u = self.u_pred[:, NUA * a:NUA * (a + 1)]
                for name, col in (("x", 0), ("y", 1), ("v_long", 2), ("psi", 3)):
                    setattr(pred, name, array.array('d', q[:, col]))
                pred.u_a = array.array('d', u[:, 0])
                pred.u_steer = array.array('d', u[:, 1])
                pred.t = t
            return
        NQA = self.nqa
        L_f, L_r = self.game.L_f, self.game.L_r
        for a, pred in enumerate(self.state_input_predictions):
            q = self.q_pred[:, NQA * a:NQA * (a + 1)]
            u = self.u_pred[:, NUA * a:NUA * (a + 1)]
            # qu2prediction (dynamics_models.py:1127-1150) incl. its psidot = v*L_r*sin(...) quirk
            psidot = q[:-1, 2] * L_r * np.sin(np.arctan(np.tan(u[:, 1]) * L_f / (L_f + L_r)))
            psidot = np.append(psidot, psidot[-1])
            for name, col in (("x", 0), ("y", 1), ("v_long", 2), ("e_psi", 3), ("s", 4), ("x_tran", 5)):
                setattr(pred, name, array.array('d', q[:, col]))
            pred.psidot = array.array('d', psidot)
            pred.v_tran = array.array('d', psidot * L_r)
            pred.u_a = array.array('d', u[:, 0])
            pred.u_steer = array.array('d', u[:, 1])
            pred.t = t

    # ------------------------------------------------------------------ batched entry point
    def solve_batch(self, x0, u_ws, l_ws=None, stream: Optional[int] = None, u_prev=None) -> BatchResult:
        """Solve B instances.  ``x0`` [B, n_q], ``u_ws`` [B, n] agent-major.  ``u_prev`` [B, n_u]: previous input of every
        instance for receding-horizon use of the v2 policy (``DGSQP_v2.py:311,328``; v1 zeroes it, ``DGSQP.py:305``).

        NumPy inputs take the host path (H2D / D2H copies inside the library call); torch CUDA
        tensors take the device path (no copies, outputs are torch tensors on the same device)."""
        g = self.game
        is_torch = hasattr(x0, "is_cuda")
        if is_torch:
            return self._solve_batch_device(x0, u_ws, l_ws, stream, u_prev=u_prev)
        x0 = np.ascontiguousarray(x0, dtype=np.float64)
        u_ws = np.ascontiguousarray(u_ws, dtype=np.float64)
        B = x0.shape[0]
        if x0.shape != (B, g.n_q) or u_ws.shape != (B, g.n):
            raise RuntimeError('Batch of shape %s / %s is incompatible with required (B,%i) / (B,%i)'
                               % (x0.shape, u_ws.shape, g.n_q, g.n))
        if l_ws is not None:
            l_ws = np.ascontiguousarray(l_ws, dtype=np.float64)
            if l_ws.shape != (B, g.m):
                raise RuntimeError('Dual warm start of shape %s is incompatible with required (B,%i)' % (l_ws.shape, g.m))
        if u_prev is not None:
            u_prev = np.ascontiguousarray(u_prev, dtype=np.float64)
            if u_prev.shape != (B, g.n_u):
                raise RuntimeError('Previous inputs of shape %s are incompatible with required (B,%i)' % (u_prev.shape, g.n_u))
        u, l = np.empty((B, g.n)), np.empty((B, g.m))
        x = np.empty((B, (g.N + 1) * g.n_q))
        cost, cond = np.empty((B, g.M)), np.empty((B, 3))
        it, st, qp = (np.empty(B, dtype=np.int32) for _ in range(3))
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        t0 = time.perf_counter()
        lw = p(l_ws) if l_ws is not None else None
        tail = (p(u), p(l), p(x), p(cost), p(cond), p(it), p(st), p(qp), 0, C.c_void_p(stream or 0))
        if u_prev is None:
            _abi.check(self._lib.dgsqp_solve_batch(self._h, B, p(x0), p(u_ws), lw, *tail))
        else:
            _abi.check(self._lib.dgsqp_solve_batch_up(self._h, B, p(x0), p(u_ws), lw, p(u_prev), *tail))
        return BatchResult(u, l, x, cost, cond, it, st, qp, time.perf_counter() - t0)

    def _solve_batch_device(self, x0, u_ws, l_ws=None, stream=None, out=None, sync=True, u_prev=None):
        import torch
        g = self.game
        B = x0.shape[0]
        if tuple(x0.shape) != (B, g.n_q) or tuple(u_ws.shape) != (B, g.n):
            raise RuntimeError('Batch of shape %s / %s is incompatible with required (B,%i) / (B,%i)'
                               % (tuple(x0.shape), tuple(u_ws.shape), g.n_q, g.n))
        if not (x0.is_cuda and u_ws.is_cuda and x0.dtype == torch.float64 and u_ws.dtype == torch.float64):
            raise RuntimeError("device path needs float64 CUDA tensors")
        x0, u_ws = x0.contiguous(), u_ws.contiguous()
        dev = x0.device
        if out is None:
            out = self.alloc_outputs(B, dev)
        u, l, x, cost, cond, it, st, qp = out
        if stream is None:
            stream = torch.cuda.current_stream(dev).cuda_stream
        v = lambda t: C.c_void_p(t.data_ptr())
        for name, t, width in (("l_ws", l_ws, g.m), ("u_prev", u_prev, g.n_u)):
            if t is not None and not (t.is_cuda and t.dtype == torch.float64 and tuple(t.shape) == (B, width) and t.device == dev):
                raise RuntimeError(f"{name} must be a float64 CUDA tensor of shape (B,{width}) on {dev}")
        # contiguous copies stay referenced by the result until the caller drops it (the kernel may still be queued)
        l_ws = l_ws.contiguous() if l_ws is not None else None
        u_prev = u_prev.contiguous() if u_prev is not None else None
        lw = v(l_ws) if l_ws is not None else None
        if u_prev is not None and not sync:
            raise RuntimeError("u_prev needs the synchronous device call")
        upv = v(u_prev) if u_prev is not None else None
        t0 = time.perf_counter()
        if sync and u_prev is not None:
            _abi.check(self._lib.dgsqp_solve_batch_up(self._h, B, v(x0), v(u_ws), lw, upv, v(u), v(l), v(x), v(cost),
                                                      v(cond), v(it), v(st), v(qp), 1, C.c_void_p(stream)))
        elif sync:
            _abi.check(self._lib.dgsqp_solve_batch(self._h, B, v(x0), v(u_ws), lw, v(u), v(l), v(x), v(cost), v(cond),
                                                   v(it), v(st), v(qp), 1, C.c_void_p(stream)))
        else:
            _abi.check(self._lib.dgsqp_solve_batch_async(self._h, B, v(x0), v(u_ws), lw, v(u), v(l), v(x), v(cost),
                                                         v(cond), v(it), v(st), v(qp), C.c_void_p(stream)))
        res = BatchResult(u, l, x, cost, cond, it, st, qp, time.perf_counter() - t0)
        res._inputs = (x0, u_ws, l_ws, u_prev)          # keeps the device buffers of an asynchronous launch alive
        return res

    def alloc_outputs(self, B, dev):
        import torch
        g = self.game
        f = dict(dtype=torch.float64, device=dev)
        i = dict(dtype=torch.int32, device=dev)
        return (torch.empty((B, g.n), **f), torch.empty((B, g.m), **f), torch.empty((B, (g.N + 1) * g.n_q), **f),
                torch.empty((B, g.M), **f), torch.empty((B, 3), **f), torch.empty(B, **i), torch.empty(B, **i),
                torch.empty(B, **i))

    def batch_stats(self, res: BatchResult, stream: Optional[int] = None) -> np.ndarray:
        """Additive statistics vector (``dgsqp_b200.sharding.STAT_KEYS``) of a device-path result, reduced on the GPU:
        only 16 doubles cross the bus (the table ``scripts/process_data_curve.py:98-110`` prints is built from it)."""
        if not hasattr(res.status, "is_cuda"):
            from .sharding import shard_stats
            return shard_stats(res.status, res.num_iters, res.qp_solves, res.cond)
        import torch
        out = np.zeros(16)
        v = lambda t: C.c_void_p(t.data_ptr())
        if stream is None:
            stream = torch.cuda.current_stream(res.status.device).cuda_stream
        _abi.check(self._lib.dgsqp_batch_stats(self.device, int(res.status.shape[0]), v(res.status), v(res.num_iters),
                                               v(res.qp_solves), v(res.cond), out.ctypes.data_as(C.c_void_p),
                                               C.c_void_p(stream)))
        return out

    def last_stats(self) -> np.ndarray:
        """The same 16-entry statistics vector for the LAST ``solve_batch`` of this solver, accumulated by the solve kernel
        itself in its epilogue (``dgsqp_last_stats``): no second pass over the outputs, no extra launch."""
        out = np.zeros(16)
        _abi.check(self._lib.dgsqp_last_stats(self._h, out.ctypes.data_as(C.c_void_p)))
        return out

    def last_iter_data(self, B):
        """Per-iteration records of the last solve_batch (needs ``params.save_iter_data``): for each of the first B instances
        the list the reference keeps as ``iter_data`` (DGSQP.py:445-452; v2: the IterationData fields of the same name,
        DGSQP_v2.py:31-52) -- ``cond`` (p_feas, comp, stat at the start of the iteration), ``qp_solves`` and ``it_time``
        of every completed iteration.  Iterates (``u_sol``, ``l_sol``) are not kept on the device."""
        if not self.save_iter_data:
            raise RuntimeError("params.save_iter_data is False: no per-iteration records were kept")
        cap = self._lib.dgsqp_iter_log_capacity(self._h)
        buf = np.zeros((B, cap, 5))
        _abi.check(self._lib.dgsqp_last_iter_data(self._h, B, buf.ctypes.data_as(C.c_void_p)))
        out = []
        for b in range(B):
            rows = buf[b][buf[b, :, 3] > 0]
            out.append([dict(cond=dict(p_feas=float(r[0]), comp=float(r[1]), stat=float(r[2])), qp_solves=int(r[3]),
                             it_time=float(r[4])) for r in rows])
        return out

    def last_diag(self, B):
        """[B, 4] int32: full evaluations, gradient-only evaluations, QP active-set iterations,
        max number of negative Hessian eigenvalues -- of the last solve_batch."""
        d = np.empty((B, _abi.NDIAG), dtype=np.int32)
        _abi.check(self._lib.dgsqp_last_diag(self._h, B, d.ctypes.data_as(C.c_void_p)))
        return d

    def last_phase_cycles(self, B):
        """[B, len(_abi.PHASES)] int64 SM-clock cycles spent per solver phase -- of the last solve_batch."""
        k = self._lib.dgsqp_phase_count()
        assert k == len(_abi.PHASES)
        d = np.empty((B, k), dtype=np.int64)
        _abi.check(self._lib.dgsqp_last_phase_cycles(self._h, B, d.ctypes.data_as(C.c_void_p)))
        return d

    def memory_plan(self):
        """dict(smem_bytes, gmem_bytes, mats_in_smem, sens_in_smem) of one CTA."""
        o = (C.c_int64 * 4)()
        _abi.check(self._lib.dgsqp_memory_plan(self._h, o))
        return dict(smem_bytes=int(o[0]), gmem_bytes=int(o[1]), mats_in_smem=bool(o[2]), sens_in_smem=bool(o[3] & 1), hot_in_smem=bool(o[3] & 2),
                    matA_in_smem=bool(o[3] & 4))

    def set_smem_limit(self, nbytes):
        _abi.check(self._lib.dgsqp_set_smem_limit(self._h, int(nbytes)))

    def configure(self, ctas_per_sm=0, threads=0):
        _abi.check(self._lib.dgsqp_configure(self._h, int(ctas_per_sm), int(threads)))
