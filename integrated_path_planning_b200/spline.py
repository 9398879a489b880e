"""Reference path: natural cubic spline over chord length.

Host-side construction only (once per planner); the device evaluates it.  Same public surface
as reference `src/planning/cubic_spline.py` (`CubicSpline1D`, `CubicSpline2D`, `calc_*`), and the
same linear system, so coefficients are bit-identical to the reference's for the same waypoints.
"""
from __future__ import annotations

import numpy as np


class CubicSpline1D:
    """Piecewise cubic y(x) with natural end conditions (cubic_spline.py:13-190)."""

    def __init__(self, x, y):
        self.x = np.array(x, dtype=float)
        self.y = self.a = np.array(y, dtype=float)
        self.nx = n = len(self.x)
        if n < 2:
            raise ValueError("need at least two knots")
        h = np.diff(self.x)
        if np.any(h < 0):
            raise ValueError("x coordinates must be sorted in ascending order")
        lhs = np.zeros((n, n))
        lhs[0, 0] = lhs[-1, -1] = 1.0
        rows = np.arange(1, n - 1)
        lhs[rows, rows - 1] = h[:-1]
        lhs[rows, rows] = 2.0 * (h[:-1] + h[1:])
        lhs[rows, rows + 1] = h[1:]
        rhs = np.zeros(n)
        rhs[1:-1] = 3.0 * (self.a[2:] - self.a[1:-1]) / h[1:] - 3.0 * (self.a[1:-1] - self.a[:-2]) / h[:-1]
        self.c = np.linalg.solve(lhs, rhs)
        self.d = (self.c[1:] - self.c[:-1]) / (3.0 * h)
        self.b = (self.a[1:] - self.a[:-1]) / h - h * (2.0 * self.c[:-1] + self.c[1:]) / 3.0

    def _locate(self, x):
        x = np.atleast_1d(np.asarray(x, dtype=float))
        inside = (x >= self.x[0]) & (x <= self.x[-1])
        seg = np.clip(np.searchsorted(self.x, x, side="right") - 1, 0, self.nx - 2)
        return x, inside, seg

    def _eval(self, x, order):
        x, inside, seg = self._locate(x)
        out = np.full(x.shape, np.nan)
        i = seg[inside]
        dx = x[inside] - self.x[i]
        if order == 0:
            out[inside] = self.a[i] + self.b[i] * dx + self.c[i] * dx ** 2.0 + self.d[i] * dx ** 3.0
        elif order == 1:
            out[inside] = self.b[i] + 2.0 * self.c[i] * dx + 3.0 * self.d[i] * dx ** 2.0
        elif order == 2:
            out[inside] = 2.0 * self.c[i] + 6.0 * self.d[i] * dx
        else:
            out[inside] = 6.0 * self.d[i]
        return out[0] if out.shape == (1,) else out

    def calc_position(self, x):
        return self._eval(x, 0)

    def calc_first_derivative(self, x):
        return self._eval(x, 1)

    def calc_second_derivative(self, x):
        return self._eval(x, 2)

    def calc_third_derivative(self, x):
        return self._eval(x, 3)


class CubicSpline2D:
    """Planar path (x(s), y(s)), s = cumulative chord length (cubic_spline.py:192-288)."""

    def __init__(self, x, y):
        x = np.asarray(x, dtype=float)
        y = np.asarray(y, dtype=float)
        self.ds = np.hypot(np.diff(x), np.diff(y))
        self.s = [0]
        self.s.extend(np.cumsum(self.ds))
        self.sx = CubicSpline1D(self.s, x)
        self.sy = CubicSpline1D(self.s, y)

    def calc_position(self, s):
        return self.sx.calc_position(s), self.sy.calc_position(s)

    def calc_yaw(self, s):
        return np.arctan2(self.sy.calc_first_derivative(s), self.sx.calc_first_derivative(s))

    def calc_curvature(self, s):
        dx, ddx = self.sx.calc_first_derivative(s), self.sx.calc_second_derivative(s)
        dy, ddy = self.sy.calc_first_derivative(s), self.sy.calc_second_derivative(s)
        return (ddy * dx - ddx * dy) / ((dx ** 2 + dy ** 2) ** (3 / 2))

    def calc_curvature_rate(self, s):
        dx, dy = self.sx.calc_first_derivative(s), self.sy.calc_first_derivative(s)
        ddx, ddy = self.sx.calc_second_derivative(s), self.sy.calc_second_derivative(s)
        dddx, dddy = self.sx.calc_third_derivative(s), self.sy.calc_third_derivative(s)
        num_a = dx * ddy - dy * ddx
        num_b = dx * dddy - dy * dddx
        num_c = dx * ddx + dy * ddy
        den = dx * dx + dy * dy
        return num_b / den ** 1.5 - 3.0 * num_a * num_c / den ** 2.5


def spline_tables(path):
    """Coefficient arrays of any CubicSpline2D-shaped object (ours or the reference's) as the
    contiguous float64 arrays the C ABI takes (include/fot.h fot_tables_t).  Raises TypeError for
    objects that do not carry numeric coefficients (e.g. mocks): the device evaluates the spline
    itself and cannot call back into Python."""
    try:
        out = {"knots": np.ascontiguousarray(path.sx.x, dtype=np.float64)}
        knots_y = np.ascontiguousarray(path.sy.x, dtype=np.float64)
        for axis, sp in (("x", path.sx), ("y", path.sy)):
            for name in "abcd":
                out[axis + name] = np.ascontiguousarray(getattr(sp, name), dtype=np.float64)
    except (AttributeError, TypeError, ValueError) as exc:
        raise TypeError("reference_path must expose CubicSpline2D coefficients (sx/sy with x,a,b,c,d)") from exc
    nx = out["knots"].shape[0]
    if nx < 2 or not np.array_equal(out["knots"], knots_y):
        raise TypeError("reference_path: x and y splines must share their knots")
    for axis in "xy":
        if out[axis + "a"].shape != (nx,) or out[axis + "c"].shape != (nx,) or \
           out[axis + "b"].shape != (nx - 1,) or out[axis + "d"].shape != (nx - 1,):
            raise TypeError("reference_path: unexpected coefficient shapes")
    return out
