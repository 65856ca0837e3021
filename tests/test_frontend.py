"""Reference-style constructor arguments -> game record (dgsqp_b200/frontend.py): the callables below restate the cost and
constraint expressions of scripts/DGSQP_ALGAMES_monte_carlo_chicane.py:222-330 in NumPy."""
import math

import numpy as np
import pytest

import dgsqp_b200 as dg
from dgsqp_b200.dynamics import (CasadiDecoupledMultiAgentDynamicsModel, CasadiKinematicBicycleCombined,
                                 KinematicBicycleConfig, MultiAgentModelConfig)
from dgsqp_b200.frontend import UnsupportedGameError, game_from_reference_args
from dgsqp_b200.types import (VehicleState, VehicleActuation, ParametricPose, Position, OrientationEuler, BodyLinearVelocity,
                              BodyAngularVelocity)


def _script_args(M=2, N=25, dt=0.1, r=0.4, rate=(10.0, math.pi), track=None, extra_cost=0.0):
    track = track or dg.chicane_game().track
    cfg = KinematicBicycleConfig(dt=dt, model_name="kinematic_bicycle_cl", noise=False, discretization_method="euler",
                                 wheel_dist_front=0.13, wheel_dist_rear=0.13, drag_coefficient=0.1, slip_coefficient=0.1)
    models = [CasadiKinematicBicycleCombined(0.0, cfg, track=track) for _ in range(M)]
    joint = CasadiDecoupledMultiAgentDynamicsModel(0.0, models, MultiAgentModelConfig(dt=dt, discretization_method="euler"))
    inf = np.inf
    ub = [VehicleState(x=Position(x=inf, y=inf), p=ParametricPose(s=inf, x_tran=1.0, e_psi=inf), e=OrientationEuler(psi=inf),
                       v=BodyLinearVelocity(v_long=inf, v_tran=inf), w=BodyAngularVelocity(w_psi=inf),
                       u=VehicleActuation(u_a=2.1, u_steer=0.436)) for _ in range(M)]
    lb = [VehicleState(x=Position(x=-inf, y=-inf), p=ParametricPose(s=-inf, x_tran=-1.0, e_psi=-inf), e=OrientationEuler(psi=-inf),
                       v=BodyLinearVelocity(v_long=-inf, v_tran=-inf), w=BodyAngularVelocity(w_psi=-inf),
                       u=VehicleActuation(u_a=-2.1, u_steer=-0.436)) for _ in range(M)]
    costs, agent_c = [], []
    for a in range(M):
        def stage(q, u, um, a=a):
            return 0.5 * (1.0 * u[0] ** 2 + 1.0 * u[1] ** 2) + 0.5 * (1.0 * (u[0] - um[0]) ** 2 + 1.0 * (u[1] - um[1]) ** 2) \
                + extra_cost * q[a * 6 + 5] ** 2

        def term(q, a=a):
            s = q[a * 6 + 4]
            return -10.0 * s + sum(5.0 * math.atan(q[b * 6 + 4] - s) for b in range(M) if b != a)

        def rate_rows(q, u, um):
            return np.array([(u[0] - um[0]) - dt * rate[0], dt * -rate[0] - (u[0] - um[0]),
                             (u[1] - um[1]) - dt * rate[1], dt * -rate[1] - (u[1] - um[1])])
        costs.append([stage] * N + [term])
        agent_c.append([rate_rows] * N + [None])
    pairs = [(a, b) for a in range(M) for b in range(a + 1, M)]

    def coll(q, *_):
        return np.array([(2 * r) ** 2 - ((q[a * 6] - q[b * 6]) ** 2 + (q[a * 6 + 1] - q[b * 6 + 1]) ** 2) for a, b in pairs])
    shared = [None] + [coll] * N
    return joint, costs, agent_c, shared, dict(ub=ub, lb=lb)


@pytest.mark.parametrize("M", [2, 3])
def test_reference_args_identify_the_game(M):
    args = _script_args(M=M)
    g = game_from_reference_args(*args, dg.chicane_params())
    ref = dg.chicane_game() if M == 2 else dg.agents_game(M, 90.0, 25)
    assert (g.M, g.N, g.dt) == (M, 25, 0.1)
    for f in ("L_f", "L_r", "c_dr", "c_da", "c_s", "mass", "half_width"):
        assert getattr(g, f) == pytest.approx(getattr(ref, f), abs=1e-12)
    for f in ("input_weight", "rate_weight", "comp_weights", "u_ub", "u_lb", "rate_ub", "rate_lb", "obs_r"):
        assert np.allclose(getattr(g, f), getattr(ref, f), atol=1e-9), f
    if M == 2:
        a, b = g.to_struct(), ref.to_struct()
        assert bytes(a) == bytes(b) or all(np.allclose(getattr(a, n), getattr(b, n)) if hasattr(getattr(a, n), "__len__")
                                           else abs(getattr(a, n) - getattr(b, n)) < 1e-9 for n, _ in a._fields_)


def test_outside_the_family_raises():
    with pytest.raises(UnsupportedGameError, match="stage cost"):
        game_from_reference_args(*_script_args(extra_cost=0.3))
    joint, costs, agent_c, shared, bounds = _script_args()
    with pytest.raises(UnsupportedGameError, match="shared constraint"):
        game_from_reference_args(joint, costs, agent_c, [None] + [lambda q, *_: np.array([q[0] - 1.0])] * 25, bounds)
    with pytest.raises(UnsupportedGameError, match="terminal agent constraints"):
        game_from_reference_args(joint, costs, [c[:-1] + [lambda q: q[:1]] for c in agent_c], shared, bounds)
    bounds["ub"][0].v.v_long = 3.0
    with pytest.raises(UnsupportedGameError, match="finite bound"):
        game_from_reference_args(joint, costs, agent_c, shared, bounds)
