"""Game descriptors: what the reference scripts assemble as CasADi objects before constructing the
solver (vehicle models, joint model, per-stage cost / constraint Functions, bounds), collected into
one record that the C ABI consumes (``dgsqp_racing_game`` in ``include/dgsqp_b200.h``).

Factories reproduce the literals of the BASELINE configurations:
``chicane_game``  scripts/DGSQP_ALGAMES_monte_carlo_chicane.py:49-174
``curve_game``    scripts/DGSQP_ALGAMES_monte_carlo_curve.py (theta, N sweep; reg = 0, steer rate 4.5, r = 0.2)
``agents_game``   scripts/DGSQP_monte_carlo_agents.py:47-153 (M agents, 90 degree curve, r = 0.4)
``merge_game``    scripts/DGSQP_merge_monte_carlo.py:40-192 (three unicycles, one on the ramp; ``dgsqp_merge_game``)
"""
from dataclasses import dataclass, field
import math
from typing import List, Tuple

import numpy as np

from ._abi import RacingGameStruct, MergeGameStruct, MAX_AGENTS, MAX_TRACK_SEGS
from .solver_types import DGSQPParams
from .tracks import RadiusArclengthTrack, ChicaneTrack, CurveTrack
from .types import VehicleState

NQA, NUA = 6, 2     # q = [x, y, v_long, e_psi, s, x_tran], u = [u_a, u_steer]


@dataclass
class RacingGame:
    track: RadiusArclengthTrack
    M: int = 2
    N: int = 25
    dt: float = 0.1
    L_f: float = 0.13
    L_r: float = 0.13
    c_dr: float = 0.1
    c_da: float = 0.0
    c_s: float = 0.1
    mass: float = 2.366
    input_weight: Tuple[float, float] = (1.0, 1.0)
    rate_weight: Tuple[float, float] = (1.0, 1.0)
    comp_weights: Tuple[float, float] = (10.0, 5.0)
    u_ub: Tuple[float, float] = (2.1, 0.436)
    u_lb: Tuple[float, float] = (-2.1, -0.436)
    rate_ub: Tuple[float, float] = (10.0, math.pi)
    rate_lb: Tuple[float, float] = (-10.0, -math.pi)
    half_width: float = 1.0
    obs_r: List[float] = field(default_factory=lambda: [0.4, 0.4])
    name: str = "racing"

    def __post_init__(self):
        if not 2 <= self.M <= MAX_AGENTS:
            raise ValueError(f"racing game supports 2..{MAX_AGENTS} agents, got {self.M}")
        if len(self.obs_r) != self.M:
            raise ValueError("Number of agents: %i, but %i collision radii were provided" % (self.M, len(self.obs_r)))

    # dimensions (DGSQP.py:157-170 and the constraint assembly :730-821)
    @property
    def n_q(self):
        return NQA * self.M

    @property
    def n_u(self):
        return NUA * self.M

    @property
    def n(self):
        return self.N * self.n_u

    @property
    def n_c(self):
        P = self.M * (self.M - 1) // 2
        return [8 * self.M] + [P + 10 * self.M] * (self.N - 1) + [P + 2 * self.M]

    @property
    def m(self):
        return int(sum(self.n_c))

    def state2q(self, states: List[VehicleState]) -> np.ndarray:
        """Joint state vector (CasadiDecoupledMultiAgentDynamicsModel.state2q, dynamics_models.py:2576-2583,
        over CasadiKinematicBicycleCombined.state2q :1086-1088)."""
        return np.array([[s.x.x, s.x.y, s.v.v_long, s.p.e_psi, s.p.s, s.p.x_tran] for s in states],
                        dtype=np.float64).ravel()

    def to_struct(self) -> RacingGameStruct:
        g = RacingGameStruct()
        g.M, g.N, g.dt = self.M, self.N, self.dt
        g.L_f, g.L_r, g.c_dr, g.c_da, g.c_s, g.mass = self.L_f, self.L_r, self.c_dr, self.c_da, self.c_s, self.mass
        for i in range(2):
            g.input_weight[i], g.rate_weight[i], g.comp_weights[i] = (self.input_weight[i], self.rate_weight[i],
                                                                      self.comp_weights[i])
            g.u_ub[i], g.u_lb[i], g.rate_ub[i], g.rate_lb[i] = self.u_ub[i], self.u_lb[i], self.rate_ub[i], self.rate_lb[i]
        g.half_width = self.half_width
        for a in range(MAX_AGENTS):
            g.obs_r[a] = self.obs_r[a] if a < self.M else 0.0
        lens, curv = self.track.segment_lengths(), self.track.segment_curvatures()
        if len(lens) > MAX_TRACK_SEGS:
            raise ValueError(f"track has {len(lens)} segments, at most {MAX_TRACK_SEGS} supported")
        g.track_nseg = len(lens)
        for i in range(len(lens)):
            g.track_seg_len[i], g.track_seg_curv[i] = lens[i], curv[i]
        return g


def chicane_game(theta_deg=45.0, N=25):
    track = ChicaneTrack(enter_straight_length=1, curve1_length=4, curve1_swept_angle=theta_deg * np.pi / 180,
                         mid_straight_length=1, curve2_length=4, curve2_swept_angle=theta_deg * np.pi / 180,
                         exit_straight_length=5, width=2.0, slack=0.8, mirror=False)
    return RacingGame(track=track, M=2, N=N, obs_r=[0.4, 0.4], name=f"chicane_{theta_deg:g}_N{N}")


def chicane_params(N=25):
    return DGSQPParams(solver_name="SQGAMES", dt=0.1, N=N, reg=1e-3, nonmono_ls=True, line_search_iters=50,
                       sqp_iters=50, p_tol=1e-3, d_tol=1e-3, beta=0.01, tau=0.5)


def curve_game(theta_deg=45.0, N=25):
    track = CurveTrack(enter_straight_length=1, curve_length=8, curve_swept_angle=theta_deg * np.pi / 180,
                       exit_straight_length=5, width=2.0, slack=0.8, ccw=True)
    return RacingGame(track=track, M=2, N=N, rate_ub=(10.0, 4.5), rate_lb=(-10.0, -4.5), obs_r=[0.2, 0.2],
                      name=f"curve_{theta_deg:g}_N{N}")


def curve_params(N=25):
    return DGSQPParams(solver_name="SQGAMES", dt=0.1, N=N, reg=0.0, nonmono_ls=True, line_search_iters=50,
                       sqp_iters=50, p_tol=1e-3, d_tol=1e-3, beta=0.01, tau=0.5)


def agents_game(M=3, theta_deg=90.0, N=25):
    track = CurveTrack(enter_straight_length=1, curve_length=8, curve_swept_angle=theta_deg * np.pi / 180,
                       exit_straight_length=5, width=2.0, slack=0.8, ccw=True)
    return RacingGame(track=track, M=M, N=N, obs_r=[0.4] * M, name=f"agents_M{M}_{theta_deg:g}_N{N}")


def agents_params(N=25):
    return DGSQPParams(solver_name="DGSQP", dt=0.1, N=N, reg=1e-3, nonmono_ls=True, line_search_iters=50,
                       sqp_iters=50, p_tol=1e-3, d_tol=1e-3, beta=0.01, tau=0.5)


def merge_lanes(lw=0.3, mw=0.3, mp=1.5, th=math.pi / 12):
    """Lane half-planes of the merge scenario per car (scripts/DGSQP_merge_monte_carlo.py:40-74,316-318): rows
    ``n(x)'(p - (pt - r n(x))) <= 0`` as tuples (brk, n_lo, n_hi, pt); cars 1, 2 drive the straight lane, car 3 the ramp
    whose normals switch to the straight lane's at x6 / x7 (``ca.pw_const``)."""
    ns = (0.0, 1.0)
    nm = (-math.sin(th), math.cos(th))
    neg = lambda v: (-v[0], -v[1])
    x1, x3 = (0.0, lw), (0.0, 0.0)
    x6 = (mp + lw / math.tan(th), lw)
    x7 = (mp + mw / math.sin(th), 0.0)
    inf = float("inf")
    straight = [(inf, ns, ns, x1), (inf, neg(ns), neg(ns), x3)]
    ramp = [(x6[0], nm, ns, x6), (x7[0], neg(nm), neg(ns), x7)]
    return [straight, straight, ramp]


@dataclass
class MergeGame:
    """Merge scenario (scripts/DGSQP_merge_monte_carlo.py): M kinematic unicycles ``q = [x, y, v, psi]``,
    ``u = [F_x, w_z]`` (``CasadiKinematicUnicycle``, dynamics_models.py:306-345) under RK3 (:202-212).  The script's
    DynamicsConfig carries no mass; the DynamicsConfig default 2.366 applies (model_types.py:99)."""
    M: int = 3
    N: int = 20
    dt: float = 0.1
    mass: float = 2.366
    input_weight: Tuple[float, float] = (0.1, 0.1)
    state_weight: Tuple[float, float, float, float] = (1.0, 10.0, 1.0, 1.0)
    term_scale: float = 10.0
    goals: List[Tuple[float, float, float, float]] = field(
        default_factory=lambda: [(4.0, 0.15, 0.3, 0.0), (4.5, 0.15, 0.3, 0.0), (4.25, 0.15, 0.3, 0.0)])
    u_ub: Tuple[float, float] = (2.0, 4.5)
    u_lb: Tuple[float, float] = (-2.0, -4.5)
    v_ub: float = 2.0
    v_lb: float = -2.0
    obs_r: List[float] = field(default_factory=lambda: [0.1, 0.1, 0.1])
    lane_r: float = 0.1
    lanes: list = field(default_factory=merge_lanes)
    name: str = "merge"

    def __post_init__(self):
        if not 2 <= self.M <= MAX_AGENTS:
            raise ValueError(f"merge game supports 2..{MAX_AGENTS} agents, got {self.M}")
        if len(self.obs_r) != self.M or len(self.goals) != self.M or len(self.lanes) != self.M:
            raise ValueError("Number of agents: %i, but %i radii / %i goals / %i lane sets were provided"
                             % (self.M, len(self.obs_r), len(self.goals), len(self.lanes)))

    @property
    def n_q(self):
        return 4 * self.M

    @property
    def n_u(self):
        return 2 * self.M

    @property
    def n(self):
        return self.N * self.n_u

    @property
    def n_c(self):
        P = self.M * (self.M - 1) // 2
        return [6 * self.M] + [P + 8 * self.M] * (self.N - 1) + [P + 4 * self.M]

    @property
    def m(self):
        return int(sum(self.n_c))

    def state2q(self, states: List[VehicleState]) -> np.ndarray:
        """CasadiKinematicUnicycle.state2q (dynamics_models.py:343-345 / state2qu :340-342)."""
        return np.array([[s.x.x, s.x.y, s.v.v_long, s.e.psi] for s in states], dtype=np.float64).ravel()

    def to_struct(self) -> MergeGameStruct:
        g = MergeGameStruct()
        g.M, g.N, g.dt, g.mass = self.M, self.N, self.dt, self.mass
        for i in range(2):
            g.input_weight[i], g.u_ub[i], g.u_lb[i] = self.input_weight[i], self.u_ub[i], self.u_lb[i]
        for i in range(4):
            g.state_weight[i] = self.state_weight[i]
        g.term_scale, g.v_ub, g.v_lb, g.lane_r = self.term_scale, self.v_ub, self.v_lb, self.lane_r
        for a in range(self.M):
            g.obs_r[a] = self.obs_r[a]
            for i in range(4):
                g.goal[a][i] = self.goals[a][i]
            for j in range(2):
                brk, n_lo, n_hi, pt = self.lanes[a][j]
                g.lane[a][j].brk = brk
                for i in range(2):
                    g.lane[a][j].n_lo[i], g.lane[a][j].n_hi[i], g.lane[a][j].pt[i] = n_lo[i], n_hi[i], pt[i]
        return g


def merge_game(N=20):
    return MergeGame(N=N, name=f"merge_N{N}")


def merge_params(N=20):
    """scripts/DGSQP_merge_monte_carlo.py:178-192."""
    return DGSQPParams(solver_name="DGSQP", dt=0.1, N=N, reg=0.0, merit_function="stat_l1", nonmono_ls=True,
                       line_search_iters=50, sqp_iters=50, p_tol=1e-3, d_tol=1e-3, beta=0.01, tau=0.5)


MU_VIO_THRESH = 1e-10     # see include/dgsqp_b200.h (dgsqp_params.mu_vio_thresh) and DESIGN.md D2


def _common(p, params, mu_vio_thresh, qp_warm_start, iter_log):
    p.mu_vio_thresh = mu_vio_thresh
    tl = getattr(params, "time_limit", None)
    if tl is not None and not tl > 0:
        raise ValueError(f"time_limit must be positive or None, got {tl}")
    p.time_limit = 0.0 if tl is None else float(tl)
    p.qp_warm_start = int(bool(qp_warm_start))
    p.iter_log = int(bool(iter_log))


def params_to_struct(params, mu_vio_thresh=MU_VIO_THRESH, qp_warm_start=True, iter_log=False):
    from ._abi import ParamsStruct
    if params.merit_function not in ("stat_l1", "stat"):
        raise ValueError(f"Merit function option {params.merit_function} not recognized")
    p = ParamsStruct()
    p.reg, p.p_tol, p.d_tol, p.beta, p.tau = params.reg, params.p_tol, params.d_tol, params.beta, params.tau
    p.line_search_iters, p.sqp_iters = params.line_search_iters, params.sqp_iters
    p.nonmono_ls = int(bool(params.nonmono_ls))
    p.merit_function = 0 if params.merit_function == "stat_l1" else 1
    p.conv_approx = int(bool(params.conv_approx))
    _common(p, params, mu_vio_thresh, qp_warm_start, iter_log)
    return p


def params_v2_to_struct(params, mu_vio_thresh=MU_VIO_THRESH, qp_warm_start=True, iter_log=False):
    """DGSQPV2Params -> dgsqp_v2_params.  Unknown option strings raise like the reference (DGSQP_v2.py:1162-1163)."""
    from ._abi import ParamsV2Struct
    if params.merit_function not in ("stat_l1", "sum_obj_l1"):
        raise ValueError(f"Merit function option {params.merit_function} not recognized")
    if params.merit_decrease_condition not in ("armijo", "max"):
        raise ValueError(f"Merit decrease condition {params.merit_decrease_condition} not recognized")
    if not 1 <= params.nms_memory_size <= 16:
        raise ValueError("nms_memory_size must be in 1..16")
    p = ParamsV2Struct()
    p.reg, p.reg_decay, p.p_tol, p.d_tol, p.beta, p.tau = (params.reg, params.reg_decay, params.p_tol, params.d_tol,
                                                          params.beta, params.tau)
    p.line_search_iters, p.sqp_iters = params.line_search_iters, params.sqp_iters
    p.nms, p.nms_frequency, p.nms_memory_size = int(bool(params.nms)), params.nms_frequency, params.nms_memory_size
    p.merit_function = 0 if params.merit_function == "stat_l1" else 1
    p.has_merit_parameter = int(params.merit_parameter is not None)
    p.merit_parameter = 0.0 if params.merit_parameter is None else float(params.merit_parameter)
    p.merit_decrease = params.merit_decrease
    p.merit_decrease_condition = 0 if params.merit_decrease_condition == "armijo" else 1
    p.delta_decay = params.delta_decay
    _common(p, params, mu_vio_thresh, qp_warm_start, iter_log)
    return p
