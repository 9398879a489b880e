"""B200-native Frenet optimal-trajectory candidate sweep (drop-in for the reference planner).

Public surface mirrors reference `src/planning/__init__.py`: `FrenetPlanner`, `CubicSpline2D`,
plus the boundary dataclasses.  The sweep runs only through the CUDA library (include/fot.h).
"""
from .types import EgoVehicleState, FrenetPath, FrenetState
from .spline import CubicSpline1D, CubicSpline2D
from .planner import FrenetPlanner
from .batch import BatchFrenetPlanner, DeviceBatch, PeerGather, WinnerBlock, gather_winners, shard_bounds

__all__ = ["FrenetPlanner", "BatchFrenetPlanner", "DeviceBatch", "CubicSpline1D", "CubicSpline2D",
           "EgoVehicleState", "FrenetPath", "FrenetState", "PeerGather", "WinnerBlock", "gather_winners", "shard_bounds"]
