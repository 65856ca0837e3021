"""Game descriptors: what the reference scripts assemble as CasADi objects before constructing the
solver (vehicle models, joint model, per-stage cost / constraint Functions, bounds), collected into
one record that the C ABI consumes (``dgsqp_racing_game`` in ``include/dgsqp_b200.h``).

Factories reproduce the literals of the BASELINE configurations:
``chicane_game``  scripts/DGSQP_ALGAMES_monte_carlo_chicane.py:49-174
``curve_game``    scripts/DGSQP_ALGAMES_monte_carlo_curve.py (theta, N sweep; reg = 0, steer rate 4.5, r = 0.2)
``agents_game``   scripts/DGSQP_monte_carlo_agents.py:47-153 (M agents, 90 degree curve, r = 0.4)
"""
from dataclasses import dataclass, field
import math
from typing import List, Tuple

import numpy as np

from ._abi import RacingGameStruct, MAX_AGENTS, MAX_TRACK_SEGS
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


MU_VIO_THRESH = 1e-10     # see include/dgsqp_b200.h (dgsqp_params.mu_vio_thresh) and DESIGN.md D2


def params_to_struct(params, mu_vio_thresh=MU_VIO_THRESH):
    from ._abi import ParamsStruct
    if params.merit_function not in ("stat_l1", "stat"):
        raise ValueError(f"Merit function option {params.merit_function} not recognized")
    p = ParamsStruct()
    p.reg, p.p_tol, p.d_tol, p.beta, p.tau = params.reg, params.p_tol, params.d_tol, params.beta, params.tau
    p.line_search_iters, p.sqp_iters = params.line_search_iters, params.sqp_iters
    p.nonmono_ls = int(bool(params.nonmono_ls))
    p.merit_function = 0 if params.merit_function == "stat_l1" else 1
    p.conv_approx = int(bool(params.conv_approx))
    p.mu_vio_thresh = mu_vio_thresh
    return p


def params_v2_to_struct(params, mu_vio_thresh=MU_VIO_THRESH):
    """DGSQPV2Params -> dgsqp_v2_params.  Unknown option strings raise like the reference (DGSQP_v2.py:1162-1163)."""
    from ._abi import ParamsV2Struct
    if params.merit_function == "sum_obj_l1":
        raise NotImplementedError("merit function 'sum_obj_l1' needs the gradient of the summed costs, which the "
                                  "condensed game evaluation does not produce; use 'stat_l1'")
    if params.merit_function != "stat_l1":
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
    p.merit_function = 0
    p.has_merit_parameter = int(params.merit_parameter is not None)
    p.merit_parameter = 0.0 if params.merit_parameter is None else float(params.merit_parameter)
    p.merit_decrease = params.merit_decrease
    p.merit_decrease_condition = 0 if params.merit_decrease_condition == "armijo" else 1
    p.delta_decay = params.delta_decay
    p.mu_vio_thresh = mu_vio_thresh
    return p
