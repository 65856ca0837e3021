"""Boundary message types: field-for-field compatible with the reference's ``DGSQP/types.py``
(``PythonMsg:14-84``, ``VehicleState:377-435``, ``VehicleActuation:170-178``,
``VehiclePrediction:484-527`` and the small pose / velocity records they nest), so that the
reference's Monte-Carlo drivers can build states and bounds unchanged.  Only the fields are
mirrored; plotting and covariance helpers are out of scope.
"""
from __future__ import annotations

import array
import copy
from dataclasses import dataclass, field, fields


@dataclass
class PythonMsg:
    """Dataclass base whose instances refuse attributes that are not declared fields."""

    def __setattr__(self, key, value):
        if key not in self.__dataclass_fields__:
            raise TypeError('Cannot add new field "%s" to frozen class %s' % (key, self))
        object.__setattr__(self, key, value)

    def copy(self):
        return copy.deepcopy(self)


def _record(name, spec, doc=""):
    """Build a PythonMsg dataclass of float fields defaulting to 0 (or the given default)."""
    ns = {"__annotations__": {}, "__doc__": doc}
    for item in spec:
        fname, default = item if isinstance(item, tuple) else (item, 0)
        ns["__annotations__"][fname] = float
        ns[fname] = field(default=default)
    return dataclass(type(name, (PythonMsg,), ns))


Position = _record("Position", ["x", "y", "z"])
VehicleActuation = _record("VehicleActuation", ["t", "u_a", "u_steer", "u_ds"])
BodyLinearVelocity = _record("BodyLinearVelocity", ["v_long", "v_tran", "v_n"])
BodyAngularVelocity = _record("BodyAngularVelocity", ["w_phi", "w_theta", "w_psi"])
BodyLinearAcceleration = _record("BodyLinearAcceleration", ["a_long", "a_tran", "a_n"])
BodyAngularAcceleration = _record("BodyAngularAcceleration", ["a_phi", "a_theta", "a_psi"])
OrientationEuler = _record("OrientationEuler", ["phi", "theta", "psi"])
OrientationQuaternion = _record("OrientationQuaternion", [("qr", 1), "qi", "qj", "qk"])
ParametricPose = _record("ParametricPose", ["s", "x_tran", "n", "e_psi"])
ParametricVelocity = _record("ParametricVelocity", ["ds", "dx_tran", "dn", "de_psi"])

_NESTED = dict(x=Position, v=BodyLinearVelocity, w=BodyAngularVelocity, a=BodyLinearAcceleration,
               aa=BodyAngularAcceleration, q=OrientationQuaternion, e=OrientationEuler, p=ParametricPose,
               pt=ParametricVelocity, u=VehicleActuation)


@dataclass
class VehicleState(PythonMsg):
    """Vehicle state: global pose ``x``/``e``, body velocities ``v``/``w``, Frenet pose ``p``,
    actuation ``u`` (reference ``DGSQP/types.py:377-435``)."""
    t: float = field(default=None)
    x: Position = field(default=None)
    v: BodyLinearVelocity = field(default=None)
    w: BodyAngularVelocity = field(default=None)
    a: BodyLinearAcceleration = field(default=None)
    aa: BodyAngularAcceleration = field(default=None)
    q: OrientationQuaternion = field(default=None)
    e: OrientationEuler = field(default=None)
    p: ParametricPose = field(default=None)
    pt: ParametricVelocity = field(default=None)
    u: VehicleActuation = field(default=None)
    du: VehicleActuation = field(default=None)
    lap_num: int = field(default=None)

    def __post_init__(self):
        for name, cls in _NESTED.items():
            if getattr(self, name) is None:
                object.__setattr__(self, name, cls())


_PRED_FIELDS = ["x", "y", "v_x", "v_y", "a_x", "a_y", "psi", "psidot", "v_long", "v_tran", "a_long", "a_tran",
                "e_psi", "s", "x_tran", "u_a", "u_steer", "u_ds"]


def _prediction_cls():
    ns = {"__annotations__": {"t": float}, "t": field(default=None),
          "__doc__": "Predicted trajectories as array.array('d') per signal (reference DGSQP/types.py:484-527)."}
    for f in _PRED_FIELDS:
        ns["__annotations__"][f] = array.array
        ns[f] = field(default=None)
    ns["__annotations__"]["lap_num"] = int
    ns["lap_num"] = field(default=None)
    return dataclass(type("VehiclePrediction", (PythonMsg,), ns))


VehiclePrediction = _prediction_cls()
