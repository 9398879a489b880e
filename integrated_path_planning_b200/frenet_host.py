"""Ego state -> Frenet initial conditions, on the host (O(1) per query, stateful).

Restates reference `src/core/coordinate_converter.py:26-88` (Cartesian->Frenet formulas) and
`:202-339` (nearest point on the path with the `_prev_s` window cache), plus the spatial->time
derivative conversion of `frenet_planner.py:362-371`.  These produce the six numbers that seed
every candidate polynomial, so they follow the reference's arithmetic literally.
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import numpy as np


def _f(v):
    """numpy 0-d / 1-element result -> scalar."""
    v = np.asarray(v)
    return v.reshape(-1)[0] if v.size == 1 else v


class CoordinateConverter:
    """Nearest-point search with the reference's cache semantics (coordinate_converter.py:185-339)."""

    WINDOW = 10.0        # half width of the cached local window [m]      (:223-224)
    WINDOW_SAMPLES = 100  # (:225)
    GLOBAL_STEP = 0.1    # (:322)

    def __init__(self, reference_path):
        self.reference_path = reference_path
        # Batched position evaluation for the ~160 points of one search: ONE calc_position call on
        # an array instead of one call per point.  NumPy evaluates array elements independently with
        # the same inner loops (including its vectorised pow), so the values are bit-identical to the
        # reference's point-by-point calls; only the per-call overhead goes (10 ms -> ~1 ms per plan()).
        self._batched = True
        # Paths that expose their coefficient arrays (ours and the reference's CubicSpline2D) get a lean
        # evaluator: the same element-wise NumPy expression as CubicSpline1D.calc_position
        # (a + b dx + c dx**2.0 + d dx**3.0), with the segment search and the powers shared by x and y.
        self._tab = None
        try:
            from .spline import spline_tables
            t = spline_tables(reference_path)
            self._tab = tuple(t[k] for k in ("knots", "xa", "xb", "xc", "xd", "ya", "yb", "yc", "yd"))
        except Exception:
            self._tab = None

    # -- search --------------------------------------------------------------------------
    def _xy(self, s):
        px, py = self.reference_path.calc_position(s)
        return _f(px), _f(py)

    def _xy_many(self, s_values):
        """Positions of several arc lengths, element for element what _xy gives."""
        if not self._batched:
            pts = [self._xy(s) for s in s_values]
            return [p[0] for p in pts], [p[1] for p in pts]
        sv = np.asarray(s_values, dtype=np.float64).reshape(-1)
        if self._tab is None:
            px, py = self.reference_path.calc_position(sv)
            return np.atleast_1d(px), np.atleast_1d(py)
        knots, xa, xb, xc, xd, ya, yb, yc, yd = self._tab
        inside = (sv >= knots[0]) & (sv <= knots[-1])                 # cubic_spline.py:62
        i = np.searchsorted(knots, sv, side="right") - 1               # cubic_spline.py:162-165
        i = np.minimum(np.maximum(i, 0), knots.shape[0] - 2)
        dx = sv - knots[i]
        dx2, dx3 = dx ** 2.0, dx ** 3.0
        px = xa[i] + xb[i] * dx + xc[i] * dx2 + xd[i] * dx3            # cubic_spline.py:73-74
        py = ya[i] + yb[i] * dx + yc[i] * dx2 + yd[i] * dx3
        if not inside.all():
            px = np.where(inside, px, np.nan)
            py = np.where(inside, py, np.nan)
        return px, py

    def _heading_curvature(self, rs):
        """yaw, curvature, curvature rate at rs: CubicSpline2D.calc_yaw / calc_curvature /
        calc_curvature_rate (cubic_spline.py:228-288) with the segment search done once.  The
        derivatives are one-element array expressions and the combinations scalar expressions,
        exactly as those methods evaluate them for a scalar argument."""
        if self._tab is None:
            return (_f(self.reference_path.calc_yaw(rs)), _f(self.reference_path.calc_curvature(rs)),
                    _f(self.reference_path.calc_curvature_rate(rs)))
        knots, xa, xb, xc, xd, ya, yb, yc, yd = self._tab
        sv = np.atleast_1d(np.asarray(rs, dtype=np.float64))
        if not ((sv >= knots[0]) & (sv <= knots[-1])).all():
            return np.nan, np.nan, np.nan
        i = np.searchsorted(knots, sv, side="right") - 1
        i = np.minimum(np.maximum(i, 0), knots.shape[0] - 2)
        h = sv - knots[i]
        h2 = h ** 2.0
        dx = (xb[i] + 2.0 * xc[i] * h + 3.0 * xd[i] * h2)[0]          # cubic_spline.py:100
        dy = (yb[i] + 2.0 * yc[i] * h + 3.0 * yd[i] * h2)[0]
        ddx = (2.0 * xc[i] + 6.0 * xd[i] * h)[0]                       # cubic_spline.py:125
        ddy = (2.0 * yc[i] + 6.0 * yd[i] * h)[0]
        dddx = (6.0 * xd[i])[0]                                        # cubic_spline.py:149
        dddy = (6.0 * yd[i])[0]
        yaw = np.arctan2(dy, dx)                                       # :287
        kappa = (ddy * dx - ddx * dy) / ((dx ** 2 + dy ** 2) ** (3 / 2))   # :246
        a = dx * ddy - dy * ddx
        b = dx * dddy - dy * dddx
        c = dx * ddx + dy * ddy
        d = dx * dx + dy * dy
        return yaw, kappa, b / d ** 1.5 - 3.0 * a * c / d ** 2.5      # :273

    def _global_search(self, x, y):
        length = self.reference_path.s[-1]
        n = max(100, int(length / self.GLOBAL_STEP))
        grid = np.linspace(0, length, n)
        px, py = self.reference_path.calc_position(grid)
        return grid[np.argmin(np.hypot(x - px, y - py))]

    def find_nearest_point_on_path(self, x, y):
        path_end = self.reference_path.s[-1]
        best_s = 0.0
        if hasattr(self, "_prev_s"):
            lo = max(0.0, self._prev_s - self.WINDOW)
            hi = min(path_end, self._prev_s + self.WINDOW)
            nearest = float("inf")
            grid = np.linspace(lo, hi, self.WINDOW_SAMPLES)
            gx, gy = self._xy_many(grid)
            for s, px, py in zip(grid, gx, gy):
                gap = math.hypot(x - px, y - py)
                if gap < nearest:
                    nearest, best_s = gap, s
            # a minimum on the window edge (not the path end) means the cache is stale (:241-248)
            stale = (abs(best_s - lo) < 1e-3 and lo > 0) or (abs(best_s - hi) < 1e-3 and hi < path_end)
            if stale:
                best_s = self._global_search(x, y)
        else:
            best_s = self._global_search(x, y)

        step = 0.2                                   # 20 three-point refinements (:253-280)
        for _ in range(20):
            s_lo = max(0, best_s - step)
            s_hi = min(path_end, best_s + step)
            px3, py3 = self._xy_many([s_lo, s_hi, best_s])
            gap_lo = math.hypot(x - px3[0], y - py3[0])
            gap_hi = math.hypot(x - px3[1], y - py3[1])
            gap_c = math.hypot(x - px3[2], y - py3[2])
            if gap_lo < gap_c and gap_lo < gap_hi:
                best_s = s_lo
            elif gap_hi < gap_c and gap_hi < gap_lo:
                best_s = s_hi
            else:
                step *= 0.5
        self._prev_s = best_s

        rs = best_s
        rx, ry = self._xy(rs)
        if np.any(np.isnan([rx, ry])):
            rs = self._global_search(x, y)
            rx, ry = self._xy(rs)
            if np.any(np.isnan([rx, ry])):
                raise ValueError(f"Failed to find valid reference point for position ({x:.2f}, {y:.2f})")
        rtheta, rkappa, rdkappa = self._heading_curvature(rs)
        if np.any(np.isnan([rtheta, rkappa, rdkappa])):
            raise ValueError(f"Failed to calculate reference path properties at s={rs:.2f}")
        return rs, rx, ry, rtheta, rkappa, rdkappa

    # -- formulas ------------------------------------------------------------------------
    @staticmethod
    def cartesian_to_frenet(rs, rx, ry, rtheta, rkappa, rdkappa, x, y, v, a, theta, kappa
                            ) -> Tuple[Tuple[float, float, float], Tuple[float, float, float]]:
        """Apollo-style conversion; returns (s, s', s''), (d, d', d'') with d', d'' SPATIAL
        derivatives (coordinate_converter.py:58-88)."""
        off_x = x - rx
        off_y = y - ry
        cos_r = np.cos(rtheta)
        sin_r = np.sin(rtheta)
        side = cos_r * off_y - sin_r * off_x
        d = np.copysign(np.hypot(off_x, off_y), side)
        delta = theta - rtheta
        tan_delta = np.tan(delta)
        cos_delta = np.cos(delta)
        one_minus_kd = 1 - rkappa * d
        d_prime = one_minus_kd * tan_delta
        kd_prime = rdkappa * d + rkappa * d_prime
        d_pprime = (-kd_prime * tan_delta +
                    one_minus_kd / (cos_delta * cos_delta) * (kappa * one_minus_kd / cos_delta - rkappa))
        s_dot = v * cos_delta / one_minus_kd
        delta_prime = one_minus_kd / cos_delta * kappa - rkappa
        s_ddot = (a * cos_delta - s_dot * s_dot * (d_prime * delta_prime - kd_prime)) / one_minus_kd
        return (rs, s_dot, s_ddot), (d, d_prime, d_pprime)


def ego_to_frenet(converter: CoordinateConverter, ego_state, last_kappa: float) -> Optional[np.ndarray]:
    """frenet_planner.py:334-374.  Returns [s, s_d, s_dd, d, d_d, d_dd] (time derivatives) or
    None when the conversion fails (the reference logs and returns None)."""
    try:
        ref = converter.find_nearest_point_on_path(ego_state.x, ego_state.y)
        (s, s_d, s_dd), (d, d_p, d_pp) = converter.cartesian_to_frenet(
            *ref, ego_state.x, ego_state.y, ego_state.v, ego_state.a, ego_state.yaw, last_kappa)
        d_d = d_p * s_d
        d_dd = d_pp * s_d ** 2 + d_p * s_dd
        return np.array([s, s_d, s_dd, d, d_d, d_dd], dtype=np.float64)
    except Exception:
        return None
