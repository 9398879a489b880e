"""CPU oracle for the Frenet candidate sweep -- TEST INFRASTRUCTURE ONLY.

This module is a NumPy restatement of the reference planner's hot path
(`/root/reference/src/planning/frenet_planner.py` and its two numeric helpers).
It exists to CHECK the CUDA path; it is never the thing shipped or measured as
the product.  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s
`cpu_baseline` / `--impl reference` legs may import it.  The package
`integrated_path_planning_b200` must never import anything from `oracle/`.

Parity status: PINNED.  `tests/golden/make_golden.py` runs the unmodified
reference (imported read-only from /root/reference) and this oracle on the same
seeded inputs and stores the reference's outputs as fixtures; on this
container the oracle reproduces the reference bit-for-bit (chosen index, cost,
all 15 winner arrays, per-candidate category and cost, last_check_stats).  The
reference holds no golden vectors of its own.

Every function cites the reference lines it restates.  The arithmetic is kept
in the reference's association order on purpose: the selected candidate is an
argmin over float64 costs that can differ by 3.5e-7 relative.
"""

from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import numpy as np

# --- constants restated from frenet_planner.py:63-91, 143-147 -----------------
LOW_SPEED_GATE = 0.5          # :63
LAT_SLIP_RATIO = 1.5          # :64
LAT_SLIP_FLOOR = 0.02         # :65
DYAW_CAP = 0.1                # :66
BRAKE_T_MIN = 0.5             # :78
BRAKE_T_STEP = 0.5            # :79
BRAKE_MIN_SPEED = 0.1         # :80
STOP_SPEED_EPS = 0.15         # :84
SINGULARITY_EPS = 0.05        # :143
EPS_S_DOT = 1e-3              # :147

# category codes shared with the CUDA path (include/fot.h)
CAT_OK, CAT_SPEED, CAT_ACCEL, CAT_CURV, CAT_LAT, CAT_ROAD, CAT_COLL, CAT_STOP, CAT_DROP = range(9)
CAT_NAMES = ("ok", "max_speed_error", "max_accel_error", "max_curvature_error",
             "max_lat_accel_error", "road_bound_error", "collision_error",
             "stop_distance_error")


# =============================================================================
# Natural cubic spline (cubic_spline.py:13-288)
# =============================================================================
class Spline1D:
    """cubic_spline.py:23-45 (construction), :47-166 (evaluation)."""

    def __init__(self, knots, values):
        self.x = np.array(knots, dtype=float)
        self.a = np.array(values, dtype=float)
        n = self.nx = len(self.x)
        h = np.diff(self.x)
        if np.any(h < 0):
            raise ValueError("knots must ascend")
        # tridiagonal system for c (cubic_spline.py:168-190)
        A = np.zeros((n, n))
        A[0, 0] = 1.0
        A[n - 1, n - 1] = 1.0
        for r in range(1, n - 1):
            A[r, r - 1] = h[r - 1]
            A[r, r] = 2.0 * (h[r - 1] + h[r])
            A[r, r + 1] = h[r]
        B = np.zeros(n)
        B[1:-1] = 3.0 * (self.a[2:] - self.a[1:-1]) / h[1:] - 3.0 * (self.a[1:-1] - self.a[:-2]) / h[:-1]
        self.c = np.linalg.solve(A, B)
        self.d = (self.c[1:] - self.c[:-1]) / (3.0 * h)
        self.b = (self.a[1:] - self.a[:-1]) / h - h * (2.0 * self.c[:-1] + self.c[1:]) / 3.0

    def _prep(self, s):
        s = np.atleast_1d(np.asarray(s, dtype=float))
        ok = (s >= self.x[0]) & (s <= self.x[-1])            # :62 inclusive domain
        seg = np.clip(np.searchsorted(self.x, s, side="right") - 1, 0, self.nx - 2)  # :162-165
        return s, ok, seg

    def value(self, s):
        s, ok, seg = self._prep(s)
        out = np.full(s.shape, np.nan)
        sv, i = s[ok], seg[ok]
        dx = sv - self.x[i]
        out[ok] = self.a[i] + self.b[i] * dx + self.c[i] * dx ** 2.0 + self.d[i] * dx ** 3.0   # :73-74
        return out

    def d1(self, s):
        s, ok, seg = self._prep(s)
        out = np.full(s.shape, np.nan)
        sv, i = s[ok], seg[ok]
        dx = sv - self.x[i]
        out[ok] = self.b[i] + 2.0 * self.c[i] * dx + 3.0 * self.d[i] * dx ** 2.0               # :100
        return out

    def d2(self, s):
        s, ok, seg = self._prep(s)
        out = np.full(s.shape, np.nan)
        sv, i = s[ok], seg[ok]
        dx = sv - self.x[i]
        out[ok] = 2.0 * self.c[i] + 6.0 * self.d[i] * dx                                        # :125
        return out

    def d3(self, s):
        s, ok, seg = self._prep(s)
        out = np.full(s.shape, np.nan)
        out[ok] = 6.0 * self.d[seg[ok]]                                                         # :149
        return out


class Spline2D:
    """cubic_spline.py:192-288: arc-length (chord) parameterised pair of splines."""

    def __init__(self, xs, ys):
        xs = np.asarray(xs, dtype=float)
        ys = np.asarray(ys, dtype=float)
        chord = np.hypot(np.diff(xs), np.diff(ys))                  # :207-209
        s = [0]
        s.extend(np.cumsum(chord))                                   # :210-211
        self.s = s
        self.sx = Spline1D(s, xs)
        self.sy = Spline1D(s, ys)

    @property
    def s_end(self):
        return self.sx.x[-1]

    def position(self, s):
        return self.sx.value(s), self.sy.value(s)

    def yaw(self, s):
        return np.arctan2(self.sy.d1(s), self.sx.d1(s))              # :284-287

    def curvature(self, s):
        dx, ddx = self.sx.d1(s), self.sx.d2(s)
        dy, ddy = self.sy.d1(s), self.sy.d2(s)
        return (ddy * dx - ddx * dy) / ((dx ** 2 + dy ** 2) ** (3 / 2))   # :246

    def curvature_rate(self, s):
        dx, dy = self.sx.d1(s), self.sy.d1(s)
        ddx, ddy = self.sx.d2(s), self.sy.d2(s)
        dddx, dddy = self.sx.d3(s), self.sy.d3(s)
        a = dx * ddy - dy * ddx
        b = dx * dddy - dy * dddx
        c = dx * ddx + dy * ddy
        d = dx * dx + dy * dy
        return b / d ** 1.5 - 3.0 * a * c / d ** 2.5                   # :273


def as_spline2d(path) -> "Spline2D":
    """Accept an oracle Spline2D, or any object carrying the reference's
    CubicSpline2D attributes (.s, .sx/.sy with .x .a .b .c .d)."""
    if isinstance(path, Spline2D):
        return path
    out = Spline2D.__new__(Spline2D)
    out.s = list(path.s)
    for name in ("sx", "sy"):
        src = getattr(path, name)
        sp = Spline1D.__new__(Spline1D)
        sp.x = np.asarray(src.x, dtype=float)
        sp.a = np.asarray(src.a, dtype=float)
        sp.b = np.asarray(src.b, dtype=float)
        sp.c = np.asarray(src.c, dtype=float)
        sp.d = np.asarray(src.d, dtype=float)
        sp.nx = len(sp.x)
        setattr(out, name, sp)
    return out


# =============================================================================
# Ego state -> Frenet state  (frenet_planner.py:334-374,
# coordinate_converter.py:26-88, 202-339)
# =============================================================================
def _scalar(v):
    v = np.asarray(v)
    return v.reshape(-1)[0] if v.size == 1 else v


class NearestPointSearch:
    """coordinate_converter.py:202-339; keeps the `_prev_s` cache (:283)."""

    def __init__(self, spline: Spline2D):
        self.sp = spline
        self.prev_s: Optional[float] = None

    def _pos(self, s):
        px, py = self.sp.position(s)
        return _scalar(px), _scalar(py)

    def _global(self, x, y):                                        # :318-339
        length = self.sp.s[-1]
        n = max(100, int(length / 0.1))
        samples = np.linspace(0, length, n)
        px, py = self.sp.position(samples)
        return samples[np.argmin(np.hypot(x - px, y - py))]

    def find(self, x, y):
        sp = self.sp
        s_last = sp.s[-1]
        best = 0.0
        if self.prev_s is not None:                                  # :221-248
            lo = max(0.0, self.prev_s - 10.0)
            hi = min(s_last, self.prev_s + 10.0)
            closest = float("inf")
            for s in np.linspace(lo, hi, 100):
                px, py = self._pos(s)
                dist = math.hypot(x - px, y - py)
                if dist < closest:
                    closest, best = dist, s
            hit_lo = abs(best - lo) < 1e-3 and lo > 0
            hit_hi = abs(best - hi) < 1e-3 and hi < s_last
            if hit_lo or hit_hi:
                best = self._global(x, y)
        else:
            best = self._global(x, y)

        step = 0.2                                                    # :253-280
        for _ in range(20):
            s_l = max(0, best - step)
            s_r = min(s_last, best + step)
            pxl, pyl = self._pos(s_l)
            pxr, pyr = self._pos(s_r)
            dl = math.hypot(x - pxl, y - pyl)
            dr = math.hypot(x - pxr, y - pyr)
            pxc, pyc = self._pos(best)
            dc = math.hypot(x - pxc, y - pyc)
            if dl < dc and dl < dr:
                best = s_l
            elif dr < dc and dr < dl:
                best = s_r
            else:
                step *= 0.5
        self.prev_s = best

        rx, ry = self._pos(best)
        if np.any(np.isnan([rx, ry])):                                # :289-296
            best = self._global(x, y)
            rx, ry = self._pos(best)
            if np.any(np.isnan([rx, ry])):
                raise ValueError("no valid reference point")
        rth = _scalar(sp.yaw(best))
        rk = _scalar(sp.curvature(best))
        rdk = _scalar(sp.curvature_rate(best))
        if np.any(np.isnan([rth, rk, rdk])):                          # :302-306
            raise ValueError("reference properties undefined")
        return best, rx, ry, rth, rk, rdk


def cartesian_to_frenet(rs, rx, ry, rth, rk, rdk, x, y, v, a, theta, kappa):
    """coordinate_converter.py:58-88."""
    dx = x - rx
    dy = y - ry
    c_r = np.cos(rth)
    s_r = np.sin(rth)
    cross = c_r * dy - s_r * dx
    d = np.copysign(np.hypot(dx, dy), cross)
    dth = theta - rth
    tan_d = np.tan(dth)
    cos_d = np.cos(dth)
    q = 1 - rk * d
    d_p = q * tan_d
    m = rdk * d + rk * d_p
    d_pp = (-m * tan_d + q / (cos_d * cos_d) * (kappa * q / cos_d - rk))
    s_dot = v * cos_d / q
    dth_p = q / cos_d * kappa - rk
    s_ddot = (a * cos_d - s_dot * s_dot * (d_p * dth_p - m)) / q
    return (rs, s_dot, s_ddot), (d, d_p, d_pp)


def ego_to_frenet(search: NearestPointSearch, ego_xyyva, last_kappa):
    """frenet_planner.py:334-374 -> (s, s_d, s_dd, d, d_d, d_dd) or None."""
    x, y, yaw, v, a = ego_xyyva
    try:
        rs, rx, ry, rth, rk, rdk = search.find(x, y)
        (s, s_d, s_dd), (d, d_p, d_pp) = cartesian_to_frenet(
            rs, rx, ry, rth, rk, rdk, x, y, v, a, yaw, last_kappa)
        d_d = d_p * s_d                                               # :368
        d_dd = d_pp * s_d ** 2 + d_p * s_dd                           # :369
        return (s, s_d, s_dd, d, d_d, d_dd)
    except Exception:
        return None


# =============================================================================
# Candidate generation (frenet_planner.py:376-503, 586-734)
# =============================================================================
@dataclass
class TimeTable:
    T: float
    t: np.ndarray
    t2: np.ndarray
    t3: np.ndarray
    t4: np.ndarray
    t5: np.ndarray
    inv4: np.ndarray
    inv5: np.ndarray


def time_table(T, dt) -> TimeTable:
    """frenet_planner.py:586-617."""
    n = int(round(T / dt))
    t = np.arange(n + 1) * dt
    t2 = t * t
    t3 = t2 * t
    t4 = t2 * t2
    t5 = t4 * t
    ts = float(T)
    m4 = np.array([[3.0 * ts ** 2, 4.0 * ts ** 3],
                   [6.0 * ts, 12.0 * ts ** 2]])
    m5 = np.array([[ts ** 3, ts ** 4, ts ** 5],
                   [3.0 * ts ** 2, 4.0 * ts ** 3, 5.0 * ts ** 4],
                   [6.0 * ts, 12.0 * ts ** 2, 20.0 * ts ** 3]])
    return TimeTable(T, t, t2, t3, t4, t5, np.linalg.inv(m4), np.linalg.inv(m5))


def lon_profiles(fs, targets, T, tt: TimeTable):
    """frenet_planner.py:619-658 -> s, s_d, s_dd, s_ddd each [n_v, n_t]."""
    a0, a1, a2 = fs[0], fs[1], fs[2] / 2.0
    tv = np.asarray(targets, dtype=float)
    rhs = np.column_stack([tv - a1 - 2.0 * a2 * T, np.full(tv.shape, -2.0 * a2)])
    co = rhs @ tt.inv4.T
    a3 = co[:, 0][:, None]
    a4 = co[:, 1][:, None]
    t, t2, t3, t4 = tt.t, tt.t2, tt.t3, tt.t4
    s = a0 + a1 * t + a2 * t2 + a3 * t3 + a4 * t4
    s_d = a1 + 2.0 * a2 * t + 3.0 * a3 * t2 + 4.0 * a4 * t3
    s_dd = 2.0 * a2 + 6.0 * a3 * t + 12.0 * a4 * t2
    s_ddd = 6.0 * a3 + 24.0 * a4 * t
    return s, s_d, s_dd, s_ddd


def lat_profiles(fs, offsets, T, tt: TimeTable):
    """frenet_planner.py:660-701 -> d, d_d, d_dd, d_ddd each [n_d, n_t]."""
    a0, a1, a2 = fs[3], fs[4], fs[5] / 2.0
    di = np.asarray(offsets, dtype=float)
    rhs = np.column_stack([di - a0 - a1 * T - a2 * T * T,
                           np.full(di.shape, -a1 - 2.0 * a2 * T),
                           np.full(di.shape, -2.0 * a2)])
    co = rhs @ tt.inv5.T
    a3 = co[:, 0][:, None]
    a4 = co[:, 1][:, None]
    a5 = co[:, 2][:, None]
    t, t2, t3, t4, t5 = tt.t, tt.t2, tt.t3, tt.t4, tt.t5
    d = a0 + a1 * t + a2 * t2 + a3 * t3 + a4 * t4 + a5 * t5
    d_d = a1 + 2.0 * a2 * t + 3.0 * a3 * t2 + 4.0 * a4 * t3 + 5.0 * a5 * t4
    d_dd = 2.0 * a2 + 6.0 * a3 * t + 12.0 * a4 * t2 + 20.0 * a5 * t3
    d_ddd = 6.0 * a3 + 24.0 * a4 * t + 60.0 * a5 * t2
    return d, d_d, d_dd, d_ddd


@dataclass
class Candidate:
    """One candidate trajectory (the oracle's own record, not FrenetPath)."""
    t: np.ndarray
    s: np.ndarray
    s_d: np.ndarray
    s_dd: np.ndarray
    s_ddd: np.ndarray
    d: np.ndarray
    d_d: np.ndarray
    d_dd: np.ndarray
    d_ddd: np.ndarray
    cost: float = float("inf")
    x: np.ndarray = None
    y: np.ndarray = None
    yaw: np.ndarray = None
    c: np.ndarray = None
    v: np.ndarray = None
    a: np.ndarray = None
    keep: int = 0
    category: int = CAT_DROP


@dataclass
class Knobs:
    """Constructor knobs (frenet_planner.py:149-210)."""
    max_speed: float = 50.0 / 3.6
    max_accel: float = 2.0
    max_curvature: float = 1.0
    dt: float = 0.2
    d_road_w: float = 0.5
    max_road_width: float = 7.0
    robot_radius: float = 2.0
    obstacle_radius: float = 0.3
    min_t: float = 4.0
    max_t: float = 5.0
    d_t_s: float = 5.0 / 3.6
    n_s_sample: int = 1
    max_lat_accel: float = 3.0
    k_j: float = 0.1
    k_t: float = 0.1
    k_d: float = 1.0
    k_s_dot: float = 1.0
    k_lat: float = 1.0
    k_lon: float = 1.0
    chance_epsilon: float = 0.0
    collision_margin_inflation: float = 1.0
    footprint_offsets: Optional[np.ndarray] = None   # EgoFootprint.offsets (footprint.py:66-81)
    footprint_radius: float = 0.0                    # EgoFootprint.radius


def horizon_grid(k: Knobs):
    n_ti = int((k.max_t - k.min_t) / k.dt + 1e-9)                    # :397
    return k.min_t + np.arange(n_ti + 1) * k.dt                      # :398


def speed_grid(k: Knobs, target_speed):
    n_down = int(target_speed / k.d_t_s + 1e-9)                      # :410
    tv = target_speed - np.arange(n_down + 1) * k.d_t_s              # :411
    if tv[-1] > 1e-9:
        tv = np.append(tv, 0.0)                                      # :412-413
    return tv


def lateral_grid(k: Knobs):
    n_side = int(k.max_road_width / k.d_road_w + 1e-9)               # :419
    return np.arange(-n_side, n_side + 1) * k.d_road_w               # :420


def brake_horizons(k: Knobs):
    return np.arange(BRAKE_T_MIN, k.min_t - 1e-9, BRAKE_T_STEP)      # :475


def path_cost(k: Knobs, c: Candidate, target_speed):
    """frenet_planner.py:703-734."""
    jp = np.sum(np.square(c.d_ddd))
    jd = (c.d[-1]) ** 2
    js = np.sum(np.square(c.s_ddd))
    jv = (target_speed - c.s_d[-1]) ** 2
    jt = c.t[-1]
    lat = k.k_j * jp + k.k_t * jt + k.k_d * jd
    lon = k.k_j * js + k.k_t * jt + k.k_s_dot * jv
    return k.k_lat * lat + k.k_lon * lon


def generate(k: Knobs, fs, target_speed) -> List[Candidate]:
    """frenet_planner.py:376-503; generation order T outer, v middle, d inner,
    brake ladder appended."""
    out: List[Candidate] = []
    for T in horizon_grid(k):
        tt = time_table(T, k.dt)
        tv = speed_grid(k, target_speed)
        if tv.size == 0:
            continue
        di = lateral_grid(k)
        if di.size == 0:
            continue
        S = lon_profiles(fs, tv, T, tt)
        D = lat_profiles(fs, di, T, tt)
        for iv in range(tv.shape[0]):
            for idd in range(di.shape[0]):
                c = Candidate(tt.t, S[0][iv], S[1][iv], S[2][iv], S[3][iv],
                              D[0][idd], D[1][idd], D[2][idd], D[3][idd])
                c.cost = path_cost(k, c, target_speed)
                out.append(c)
    # brake ladder (:453-503)
    if fs[1] > BRAKE_MIN_SPEED:
        n_total = int(round(k.max_t / k.dt)) + 1
        t_full = np.arange(n_total) * k.dt
        for Tb in brake_horizons(k):
            tt = time_table(Tb, k.dt)
            S = lon_profiles(fs, np.array([0.0]), Tb, tt)
            D = lat_profiles(fs, np.array([fs[3]]), Tb, tt)
            pad = n_total - len(tt.t)
            if pad < 0:
                continue
            ext = lambda arr, val: np.concatenate([arr, np.full(pad, val)])
            c = Candidate(t_full,
                          ext(S[0][0], S[0][0][-1]), ext(S[1][0], 0.0), ext(S[2][0], 0.0), ext(S[3][0], 0.0),
                          ext(D[0][0], D[0][0][-1]), ext(D[1][0], 0.0), ext(D[2][0], 0.0), ext(D[3][0], 0.0))
            c.cost = path_cost(k, c, target_speed)
            out.append(c)
    return out


# =============================================================================
# Frenet -> global (frenet_planner.py:736-889, coordinate_converter.py:91-182)
# =============================================================================
def wrap_angle(a):
    return np.angle(np.exp(1j * a))                                   # :182


def frenet_to_cartesian(rx, ry, rth, rk, rdk, s_d, s_dd, d, d_p, d_pp):
    """coordinate_converter.py:128-158."""
    c_r = np.cos(rth)
    s_r = np.sin(rth)
    x = rx - s_r * d
    y = ry + c_r * d
    q = 1 - rk * d
    tan_d = d_p / q
    dth = np.arctan2(d_p, q)
    cos_d = np.cos(dth)
    theta = wrap_angle(dth + rth)
    m = rdk * d + rk * d_p
    kappa = (((d_pp + m * tan_d) * cos_d * cos_d) / q + rk) * cos_d / q
    d_dot = d_p * s_d
    v = np.sqrt(q * q * s_d * s_d + d_dot * d_dot)
    dth_p = q / cos_d * kappa - rk
    a = (s_dd * q / cos_d + s_d * s_d / cos_d * (d_p * dth_p - m))
    return x, y, theta, kappa, v, a


def to_global(sp: Spline2D, cands: Sequence[Candidate]):
    """frenet_planner.py:751-887.  Fills x..a and `keep` (number of samples
    surviving the NaN-prefix truncation; 0 = candidate emptied)."""
    if not cands:
        return
    sizes = np.array([len(c.s) for c in cands])
    ends = np.cumsum(sizes)
    cat = lambda name: np.concatenate([getattr(c, name) for c in cands])
    S, Sd, Sdd = cat("s"), cat("s_d"), cat("s_dd")
    D, Dd, Ddd = cat("d"), cat("d_d"), cat("d_dd")
    rx, ry = sp.position(S)
    rth = sp.yaw(S)
    rk = sp.curvature(S)
    rdk = sp.curvature_rate(S)
    with np.errstate(all="ignore"):
        moving = np.abs(Sd) > EPS_S_DOT                               # :792
        safe = np.where(moving, Sd, 1.0)
        d_p = np.where(moving, Dd / safe, 0.0)
        d_pp = np.where(moving, (Ddd - d_p * Sdd) / (safe * safe), 0.0)
        x, y, th, kap, v, a = frenet_to_cartesian(rx, ry, rth, rk, rdk, Sd, Sdd, D, d_p, d_pp)
        q = 1.0 - rk * D                                              # :826
        singular = np.isfinite(q) & (q <= SINGULARITY_EPS)
    lo = 0
    for c, hi in zip(cands, ends):
        xs = x[lo:hi].copy()
        if np.any(singular[lo:hi]):
            xs[0] = np.nan                                            # :831-832
        bad = np.isnan(xs)
        n = hi - lo
        if np.any(bad):
            first = int(np.argmax(bad))
            n = first if first >= 2 else 0                            # :866
        c.keep = n
        c.x, c.y, c.yaw = xs[:n], y[lo:hi][:n], th[lo:hi][:n]
        c.c, c.v, c.a = kap[lo:hi][:n], v[lo:hi][:n], a[lo:hi][:n]
        if n != hi - lo:
            for name in ("t", "s", "s_d", "s_dd", "s_ddd", "d", "d_d", "d_dd", "d_ddd"):
                setattr(c, name, getattr(c, name)[:n])                # :873-875
        lo = hi


# =============================================================================
# Validity filter + collision (frenet_planner.py:891-1233)
# =============================================================================
def curvature_ok(c: Candidate, kmax):
    """frenet_planner.py:995-1033."""
    for i in range(1, c.keep):
        if c.v[i] > LOW_SPEED_GATE:
            if abs(c.c[i]) > kmax:
                return False
        else:
            dd = abs(c.d[i] - c.d[i - 1])
            ds_f = abs(c.s[i] - c.s[i - 1])
            if dd > max(LAT_SLIP_RATIO * ds_f, LAT_SLIP_FLOOR):
                return False
            dy_ = c.yaw[i] - c.yaw[i - 1]
            dyaw = abs(np.arctan2(np.sin(dy_), np.cos(dy_)))
            ds = float(np.hypot(c.x[i] - c.x[i - 1], c.y[i] - c.y[i - 1]))
            if dyaw > max(kmax * ds, DYAW_CAP):
                return False
    return True


def collision_geometry(k: Knobs, c: Candidate, inflation=1.0):
    """frenet_planner.py:1126-1179."""
    pts = np.stack([c.x, c.y], axis=1)
    tt = np.asarray(c.t, dtype=float)
    if k.footprint_offsets is None:
        ego_r = k.robot_radius
    else:
        ego_r = k.footprint_radius
        heading = np.stack([np.cos(c.yaw), np.sin(c.yaw)], axis=1)
        off = np.asarray(k.footprint_offsets, dtype=float)
        pts = (pts[None, :, :] + off[:, None, None] * heading[None, :, :]).reshape(len(off) * len(tt), 2)
        tt = np.tile(tt, len(off))
    r = max(ego_r + k.obstacle_radius, 1e-6)
    r_dyn = r * inflation
    pad = max(r, r_dyn)
    return pts, tt, pts.min(axis=0) - pad, pts.max(axis=0) + pad, r ** 2, r_dyn ** 2


def hits_static(pts, lo, hi, static, r2):
    """frenet_planner.py:1181-1198."""
    if static is None or len(static) == 0:
        return False
    inside = ((static[:, 0] >= lo[0]) & (static[:, 0] <= hi[0]) &
              (static[:, 1] >= lo[1]) & (static[:, 1] <= hi[1]))
    if not np.any(inside):
        return False
    diff = pts[:, None, :] - static[inside][None, :, :]
    return bool(np.any(np.sum(diff ** 2, axis=2) <= r2))


def hits_dynamic(pts, tt, lo, hi, dyn, r2, dt):
    """frenet_planner.py:1200-1233 (same-time-index test)."""
    if dyn is None or dyn.size == 0 or dyn.shape[-1] != 2:
        return False
    omin = np.min(dyn, axis=1)
    omax = np.max(dyn, axis=1)
    near = ((omax[:, 0] >= lo[0]) & (omin[:, 0] <= hi[0]) &
            (omax[:, 1] >= lo[1]) & (omin[:, 1] <= hi[1]))
    if not np.any(near):
        return False
    sel = dyn[near]
    idx = np.clip(np.round(tt / dt).astype(int), 0, sel.shape[1] - 1)
    diff = pts[:, None, :] - sel.transpose(1, 0, 2)[idx]
    return bool(np.any(np.sum(diff ** 2, axis=2) <= r2))


def collision_free(k: Knobs, c: Candidate, static, dyn, dist):
    """frenet_planner.py:1035-1124."""
    if dist is not None and dist.size > 0:
        pts, tt, lo, hi, r2, _ = collision_geometry(k, c)
        if hits_static(pts, lo, hi, static, r2):
            return False
        n = dist.shape[0]
        allowed = int(np.floor(k.chance_epsilon * n))
        bad = 0
        for j in range(n):
            if hits_dynamic(pts, tt, lo, hi, dist[j], r2, k.dt):
                bad += 1
                if bad > allowed:
                    return False
        return True
    pts, tt, lo, hi, r2, r2d = collision_geometry(k, c, k.collision_margin_inflation)
    if hits_static(pts, lo, hi, static, r2):
        return False
    if hits_dynamic(pts, tt, lo, hi, dyn, r2d, k.dt):
        return False
    return True


def classify(k: Knobs, cands, static, dyn, overrides, dist):
    """frenet_planner.py:891-993.  Sets c.category for every candidate."""
    vmax, amax, kmax, latmax = k.max_speed, k.max_accel, k.max_curvature, k.max_lat_accel
    if overrides:
        vmax = overrides.get("max_speed", vmax)
        amax = overrides.get("max_accel", amax)
        kmax = overrides.get("max_curvature", kmax)
        latmax = overrides.get("max_lat_accel", latmax)
    for c in cands:
        c.category = CAT_DROP
        n = c.keep
        if n == 0:
            continue
        if not (np.all(np.isfinite(c.v)) and np.all(np.isfinite(c.a)) and np.all(np.isfinite(c.c))):
            continue
        if n >= 2:
            step = np.hypot(np.diff(c.x), np.diff(c.y))
            if np.max(step) > max(vmax, k.max_speed) * k.dt * 3.0:   # :955
                continue
        if np.any(c.v[1:] > vmax):
            c.category = CAT_SPEED
        elif np.any(np.abs(c.a[1:]) > amax):
            c.category = CAT_ACCEL
        elif not curvature_ok(c, kmax):
            c.category = CAT_CURV
        elif np.any(c.v[1:] * c.v[1:] * np.abs(c.c[1:]) > latmax):
            c.category = CAT_LAT
        elif np.any(np.abs(c.d[1:]) > k.max_road_width + 1e-9):
            c.category = CAT_ROAD
        elif not collision_free(k, c, static, dyn, dist):
            c.category = CAT_COLL
        else:
            c.category = CAT_OK


def threshold_margins(k: Knobs, cands, overrides=None) -> Dict[str, np.ndarray]:
    """How close the checked quantities come to their limits ("parity tests should log
    threshold-margin histograms"): for every candidate that reaches a test of the priority chain
    (frenet_planner.py:964-984), the smallest relative distance |x_n - limit| / limit over its checked samples
    n >= 1.  The CUDA sweep evaluates these tests in algebraically equal, squared forms that differ from the
    reference's values by a few ulp (~1e-15 relative); a margin below that is where a category could flip.
    Call after to_global / classify.  Returns {test: margins of the candidates that evaluate it}."""
    vmax, amax, kmax, latmax = k.max_speed, k.max_accel, k.max_curvature, k.max_lat_accel
    if overrides:
        vmax = overrides.get("max_speed", vmax)
        amax = overrides.get("max_accel", amax)
        kmax = overrides.get("max_curvature", kmax)
        latmax = overrides.get("max_lat_accel", latmax)
    out = {"speed": [], "accel": [], "curvature": [], "lat_accel": [], "road": []}

    def rel(x, lim):
        return float(np.min(np.abs(x - lim))) / abs(lim) if len(x) and lim != 0 else np.inf

    for c in cands:
        if c.category == CAT_DROP or c.keep < 2:
            continue
        out["speed"].append(rel(c.v[1:], vmax))
        if c.category == CAT_SPEED:
            continue
        out["accel"].append(rel(np.abs(c.a[1:]), amax))
        if c.category == CAT_ACCEL:
            continue
        fast = c.v[1:] > LOW_SPEED_GATE
        out["curvature"].append(rel(np.abs(c.c[1:])[fast], kmax))
        if c.category == CAT_CURV:
            continue
        out["lat_accel"].append(rel(c.v[1:] * c.v[1:] * np.abs(c.c[1:]), latmax))
        if c.category == CAT_LAT:
            continue
        out["road"].append(rel(np.abs(c.d[1:]), k.max_road_width + 1e-9))
    return {name: np.asarray(v, dtype=float) for name, v in out.items()}


def stop_filter(cands, max_stop_distance):
    """frenet_planner.py:307-324."""
    for c in cands:
        if c.category != CAT_OK:
            continue
        stops = c.keep > 0 and abs(c.v[-1]) <= STOP_SPEED_EPS
        travel = float(c.s[-1] - c.s[0]) if c.keep > 0 else 0.0
        if not (stops and travel <= max_stop_distance + 1e-6):
            c.category = CAT_STOP


# =============================================================================
# plan()  (frenet_planner.py:227-304)
# =============================================================================
@dataclass
class OracleResult:
    best_index: int                     # generation-order index, -1 = no path
    cost: float
    arrays: Optional[Dict[str, np.ndarray]]   # the 15 winner sequences
    stats: Dict[str, int]
    categories: np.ndarray              # per candidate, CAT_*
    costs: np.ndarray                   # per candidate
    frenet_state: Optional[tuple] = None
    n_points: np.ndarray = None         # untruncated samples per candidate


class OraclePlanner:
    """Drop-in shaped like the reference planner, returning OracleResult."""

    def __init__(self, spline, knobs: Knobs):
        self.sp = as_spline2d(spline)
        self.k = knobs
        self.search = NearestPointSearch(self.sp)
        self.last_kappa = 0.0
        self.last_check_stats = None

    def reset_ego_curvature(self):
        self.last_kappa = 0.0                                          # :326-332

    def plan_frenet(self, fs, static, dyn=None, target_speed=30.0 / 3.6, overrides=None,
                    dist=None, max_stop_distance=None) -> OracleResult:
        k = self.k
        cands = generate(k, fs, target_speed)
        n_points = np.array([len(c.t) for c in cands], dtype=np.int64)
        to_global(self.sp, cands)
        static = None if static is None else np.asarray(static, dtype=float)
        classify(k, cands, static, dyn, overrides, dist)
        self.last_candidates = cands                                    # diagnostics (threshold_margins)
        if max_stop_distance is not None:
            stop_filter(cands, max_stop_distance)
        cats = np.array([c.category for c in cands], dtype=np.int8)
        costs = np.array([c.cost for c in cands], dtype=float)
        stats = {CAT_NAMES[j]: int(np.sum(cats == j)) for j in range(1, 7)}
        stats["ok"] = int(np.sum(cats == CAT_OK))
        if max_stop_distance is not None:
            stats["stop_distance_error"] = int(np.sum(cats == CAT_STOP))
        best, best_cost = -1, float("inf")
        for i, c in enumerate(cands):                                   # :1254-1257
            if c.category == CAT_OK and c.cost < best_cost:
                best, best_cost = i, c.cost
        arrays = None
        if best >= 0:
            w = cands[best]
            arrays = {n: np.asarray(getattr(w, n), dtype=float).copy()
                      for n in ("t", "s", "s_d", "s_dd", "s_ddd", "d", "d_d", "d_dd", "d_ddd",
                                "x", "y", "yaw", "c", "v", "a")}
        return OracleResult(best, best_cost if best >= 0 else float("inf"), arrays, stats,
                            cats, costs, tuple(fs), n_points)

    def plan(self, ego_xyyva, static, dyn=None, target_speed=30.0 / 3.6, overrides=None,
             dist=None, max_stop_distance=None) -> Optional[OracleResult]:
        self.last_check_stats = None
        fs = ego_to_frenet(self.search, ego_xyyva, self.last_kappa)
        if fs is None:
            return None
        res = self.plan_frenet(fs, static, dyn, target_speed, overrides, dist, max_stop_distance)
        self.last_check_stats = res.stats
        if res.best_index >= 0 and len(res.arrays["c"]) > 1:
            self.last_kappa = float(res.arrays["c"][1])                 # :301-302
        return res


def dense_evals(n_points: np.ndarray, n_circles: int, n_samples: int, n_peds: int, n_static: int = 0) -> int:
    """densely credited point-vs-obstacle tests of one plan()."""
    return int(n_points.sum()) * max(n_circles, 1) * (n_samples * n_peds + n_static)
