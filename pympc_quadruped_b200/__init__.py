"""B200-native batched convex-MPC engine (drop-in for pympc-quadruped's linear_mpc/mpc.py path).

Importing the package needs neither a GPU nor the built library; constructing an engine or a
controller does (there is no CPU fallback).
"""
from .configs import A1Config, AliengoConfig, LinearMpcConfig, RobotConfig, with_horizon  # noqa: F401
from .gait import BatchedGaitSchedule, Gait, GaitSchedule, gait_tables  # noqa: F401


def __getattr__(name):
    if name in ("MpcqEngine", "SolveResult"):
        from . import engine
        return getattr(engine, name)
    if name in ("BatchedModelPredictiveController", "ModelPredictiveController", "BatchedRobotData"):
        from . import controller
        return getattr(controller, name)
    if name in ("BatchedSwingFootTrajectoryGenerator", "BatchedLegController", "BatchedLegKinematics"):
        from . import legs
        return getattr(legs, name)
    raise AttributeError(name)
