"""Boundary types of the planner, API-compatible with the reference's dataclasses.

`EgoVehicleState`, `FrenetState` and `FrenetPath` carry the same fields, defaults and helper
methods as reference `src/core/data_structures.py:32-62, 119-146, 149-220`, so a returned
`FrenetPath` can be stored and consumed by the reference's `IntegratedSimulator` unchanged
(`integrated_simulator.py:660-667`: `len(path)`, `get_state_at_index(1)`, `.c[1]`).
Objects of the reference's own classes are accepted wherever these are (duck typing).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List

import numpy as np


@dataclass
class EgoVehicleState:
    """Ego state in the global frame (data_structures.py:32-62)."""
    x: float
    y: float
    yaw: float
    v: float
    a: float
    jerk: float = 0.0
    timestamp: float = 0.0
    state: object = None

    def to_array(self) -> np.ndarray:
        return np.array([self.x, self.y, self.yaw, self.v, self.a, self.jerk])

    @classmethod
    def from_array(cls, arr, timestamp: float = 0.0) -> "EgoVehicleState":
        jerk = arr[5] if len(arr) > 5 else 0.0
        return cls(x=arr[0], y=arr[1], yaw=arr[2], v=arr[3], a=arr[4], jerk=jerk, timestamp=timestamp)


@dataclass
class FrenetState:
    """State in the Frenet frame with TIME derivatives of d (data_structures.py:119-146)."""
    s: float
    s_d: float
    s_dd: float
    d: float
    d_d: float
    d_dd: float

    def to_array(self) -> np.ndarray:
        return np.array([self.s, self.s_d, self.s_dd, self.d, self.d_d, self.d_dd])

    @classmethod
    def from_array(cls, arr) -> "FrenetState":
        return cls(s=arr[0], s_d=arr[1], s_dd=arr[2], d=arr[3], d_d=arr[4], d_dd=arr[5])


@dataclass
class FrenetPath:
    """Winner trajectory (data_structures.py:149-220).  As in the reference the nine Frenet
    sequences are ndarrays and the six Cartesian ones are Python lists."""
    t: List[float] = field(default_factory=list)
    s: List[float] = field(default_factory=list)
    s_d: List[float] = field(default_factory=list)
    s_dd: List[float] = field(default_factory=list)
    s_ddd: List[float] = field(default_factory=list)
    d: List[float] = field(default_factory=list)
    d_d: List[float] = field(default_factory=list)
    d_dd: List[float] = field(default_factory=list)
    d_ddd: List[float] = field(default_factory=list)
    x: List[float] = field(default_factory=list)
    y: List[float] = field(default_factory=list)
    yaw: List[float] = field(default_factory=list)
    v: List[float] = field(default_factory=list)
    a: List[float] = field(default_factory=list)
    c: List[float] = field(default_factory=list)
    cost: float = float("inf")

    def __len__(self) -> int:
        if len(self.t) == 0:
            return 0
        return min(len(seq) for seq in (self.t, self.x, self.y, self.yaw, self.v, self.a))

    def get_state_at_index(self, idx: int) -> EgoVehicleState:
        if idx < 0 or idx >= len(self):
            raise IndexError(f"Index {idx} out of range for path of length {len(self)}")
        return EgoVehicleState(x=self.x[idx], y=self.y[idx], yaw=self.yaw[idx], v=self.v[idx],
                               a=self.a[idx], timestamp=self.t[idx])


SERIES = ("t", "s", "s_d", "s_dd", "s_ddd", "d", "d_d", "d_dd", "d_ddd", "x", "y", "yaw", "c", "v", "a")
"""Row order of the winner block returned by the C ABI (include/fot.h FOT_N_SERIES)."""
