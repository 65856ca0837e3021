"""Model descriptors: the objects the reference scripts pass to ``DGSQP.__init__`` as ``joint_dynamics``.

In the reference these are CasADi symbolic models (``DGSQP/dynamics/dynamics_models.py``: ``CasadiKinematicBicycleCombined``
:997-1150, ``CasadiKinematicUnicycle`` :306-390, ``CasadiDecoupledMultiAgentDynamicsModel`` :2482-2632) configured by the
records of ``DGSQP/dynamics/model_types.py``.  CasADi graphs cannot be evaluated on a GPU (and CasADi is not installable
offline), so here the same class names carry only what the device models need -- the configuration literals and the
track -- plus a NumPy ``fc`` / ``step`` for host-side roll-outs.  :mod:`dgsqp_b200.frontend` turns a joint model and
the scripts' cost / constraint callables into a game record.
"""
from dataclasses import field, make_dataclass

import numpy as np

from .types import PythonMsg

_MODEL = [("model_name", str, "model"), ("use_mx", bool, False), ("enable_jacobians", bool, True),
          ("compute_hessians", bool, False), ("verbose", bool, False), ("code_gen", bool, False), ("jit", bool, True),
          ("opt_flag", str, "O0"), ("install", bool, True), ("install_dir", str, "~/.dgsqp_models")]
_DYN = _MODEL + [("track_name", str, None), ("dt", float, 0.01), ("discretization_method", str, "euler"), ("M", int, 10),
                 ("noise", bool, False), ("noise_cov", np.ndarray, None)]
_GEOM = [("wheel_dist_front", float, 0.13), ("wheel_dist_rear", float, 0.13), ("wheel_dist_center_front", float, 0.1),
         ("wheel_dist_center_rear", float, 0.1), ("bump_dist_front", float, 0.15), ("bump_dist_rear", float, 0.15),
         ("bump_dist_center", float, 0.1), ("bump_dist_top", float, 0.1), ("com_height", float, 0.05)]
_POINT = [("mass", float, 2.366), ("damping_coefficient", float, 0.0), ("drag_coefficient", float, 0.0),
          ("rolling_resistance", float, 0.0), ("rolling_resistance_exponent", float, 0.5)]
_TABLES = {
    # model_types.py:21-31, 87-104, 114-125
    "DynamicsConfig": _DYN,
    "KinematicBicycleConfig": _DYN + _GEOM + [("mass", float, 2.366), ("drag_coefficient", float, 0.0),
                                              ("damping_coefficient", float, 0.0), ("slip_coefficient", float, 0.0),
                                              ("rolling_resistance", float, 0.0),
                                              ("rolling_resistance_exponent", float, 0.5)],
    "UnicycleConfig": _DYN + _POINT,
    "MultiAgentModelConfig": _DYN,
}


def _build(name):
    cls = make_dataclass(name, [(n, t, field(default=d)) for n, t, d in _TABLES[name]], bases=(PythonMsg,))
    cls.__module__ = __name__
    return cls


DynamicsConfig = _build("DynamicsConfig")
KinematicBicycleConfig = _build("KinematicBicycleConfig")
UnicycleConfig = _build("UnicycleConfig")
MultiAgentModelConfig = _build("MultiAgentModelConfig")


class CasadiKinematicBicycleCombined:
    """Kinematic bicycle in Frenet + global coordinates, ``q = [x, y, v, e_psi, s, x_tran]``, ``u = [u_a, u_steer]``
    (dynamics_models.py:997-1079).  Attribute names follow the reference (``L_f, L_r, m, c_dr, c_da, c_s, track``)."""
    n_q, n_u = 6, 2

    def __init__(self, t0, model_config, track=None):
        self.t0, self.model_config, self.track = t0, model_config, track
        self.dt = model_config.dt
        self.L_f, self.L_r = model_config.wheel_dist_front, model_config.wheel_dist_rear
        self.m = model_config.mass
        self.c_dr, self.c_da = model_config.drag_coefficient, model_config.damping_coefficient
        self.c_s = model_config.slip_coefficient
        self.c_r, self.p_r = model_config.rolling_resistance, model_config.rolling_resistance_exponent
        if self.c_r != 0.0:
            raise NotImplementedError("rolling resistance is not part of the device model (all BASELINE configs use 0)")
        if model_config.discretization_method != "euler":
            raise NotImplementedError("the device bicycle model is discretised by explicit Euler like the racing scripts")

    def state2q(self, state):
        return np.array([state.x.x, state.x.y, state.v.v_long, state.p.e_psi, state.p.s, state.p.x_tran])


class CasadiKinematicUnicycle:
    """Kinematic unicycle ``q = [x, y, v, psi]``, ``u = [F_x, w_z]`` (dynamics_models.py:306-345); the merge script passes a
    ``DynamicsConfig`` without ``mass`` -- the bicycle/unicycle default 2.366 applies."""
    n_q, n_u = 4, 2

    def __init__(self, t0, model_config, track=None):
        self.t0, self.model_config, self.track = t0, model_config, track
        self.dt = model_config.dt
        self.m = getattr(model_config, "mass", 2.366)
        if model_config.discretization_method != "rk3":
            raise NotImplementedError("the device unicycle model is discretised by RK3 with one sub-step like the merge script")

    def state2q(self, state):
        return np.array([state.x.x, state.x.y, state.v.v_long, state.e.psi])


class CasadiDecoupledMultiAgentDynamicsModel:
    """Joint model of dynamically decoupled agents (dynamics_models.py:2482-2528): ``q = [q^1; q^2; ...]``."""

    def __init__(self, t0, dynamics_models, model_config):
        self.t0, self.dynamics_models, self.model_config = t0, list(dynamics_models), model_config
        self.n_a = len(self.dynamics_models)
        self.n_q = sum(m.n_q for m in self.dynamics_models)
        self.n_u = sum(m.n_u for m in self.dynamics_models)
        self.dt = model_config.dt

    def state2q(self, states):
        return np.concatenate([m.state2q(s) for m, s in zip(self.dynamics_models, states)])
