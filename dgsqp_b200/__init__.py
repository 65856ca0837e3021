"""dgsqp_b200: batched Dynamic-Game SQP (DG-SQP) on NVIDIA B200.

Host-side mirror of the reference solver surface (zhu-edward/DGSQP ``DGSQP.solvers.DGSQP``) over a
C-ABI CUDA library; see DESIGN.md and INTEGRATION.md.
"""
from .solver_types import DGSQPParams, DGSQPV2Params, PIDParams
from .types import (VehicleState, VehicleActuation, VehiclePrediction, Position, ParametricPose, OrientationEuler,
                    BodyLinearVelocity, BodyAngularVelocity)
from .tracks import RadiusArclengthTrack, ChicaneTrack, CurveTrack, StraightTrack
from .games import (RacingGame, MergeGame, chicane_game, curve_game, agents_game, merge_game, chicane_params,
                    curve_params, agents_params, merge_params)

__all__ = ["DGSQPParams", "DGSQPV2Params", "PIDParams", "VehicleState", "VehicleActuation", "VehiclePrediction",
           "Position", "ParametricPose", "OrientationEuler", "BodyLinearVelocity", "BodyAngularVelocity",
           "RadiusArclengthTrack", "ChicaneTrack", "CurveTrack", "StraightTrack", "RacingGame", "chicane_game",
           "curve_game", "agents_game", "chicane_params", "curve_params", "agents_params", "MergeGame", "merge_game",
           "merge_params", "DGSQP"]


def __getattr__(name):
    if name == "DGSQP":
        from .solver import DGSQP
        return DGSQP
    raise AttributeError(name)
