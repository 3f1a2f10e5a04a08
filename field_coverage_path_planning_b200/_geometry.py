"""Host-side field geometry of the drop-in API (A2/A3 scalar setup; FP64 numpy).

The reference returns Shapely geometries under ``result['main_work']['area']`` /
``result['headland']['area']`` and callers read ``planner.field_polygon.centroid.coords[0]`` /
``.area`` (multi_field_planner.py:117-125).  Shapely is not a dependency of this package, so the
light objects below expose the attributes those callers touch: ``area``, ``bounds``,
``centroid`` (``.x .y .coords``), ``exterior.coords``, ``is_empty``.

Geometry decisions: D1 = ``buffer(-d)`` of a convex CCW quad is the exact mitred
inset with vertices in input order; D2 = ``buffer(+r)`` is the exact round buffer.
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence, Tuple

import numpy as np

Pt = Tuple[float, float]


class _Point:
    def __init__(self, x, y):
        self.x, self.y = float(x), float(y)
        self.coords = [(self.x, self.y)]


class _Ring:
    def __init__(self, verts):
        self.coords = [tuple(map(float, v)) for v in verts] + [tuple(map(float, verts[0]))]


def shoelace(verts: Sequence[Pt]) -> float:
    """Signed area, sequential FP64 sum (the device kernel uses the same order)."""
    s = 0.0
    n = len(verts)
    for i in range(n):
        x0, y0 = verts[i]
        x1, y1 = verts[(i + 1) % n]
        s += x0 * y1 - x1 * y0
    return 0.5 * s


def poly_centroid(verts: Sequence[Pt]) -> Pt:
    """Area centroid, coordinates shifted to vertex 0 first (fields may sit km from the origin)."""
    n = len(verts)
    ox, oy = verts[0]
    a2 = cx = cy = 0.0
    for i in range(n):
        x0, y0 = verts[i][0] - ox, verts[i][1] - oy
        x1, y1 = verts[(i + 1) % n][0] - ox, verts[(i + 1) % n][1] - oy
        cr = x0 * y1 - x1 * y0
        a2 += cr
        cx += (x0 + x1) * cr
        cy += (y0 + y1) * cr
    return (ox + cx / (3.0 * a2), oy + cy / (3.0 * a2))


def mitred_inset(verts: Sequence[Pt], d: float) -> Optional[List[Pt]]:
    """D1 (reference: Polygon.buffer(-d), mlp3:595, :871, :965)."""
    v = np.asarray(verts, dtype=np.float64)
    n = len(v)
    e = np.roll(v, -1, axis=0) - v
    ln = np.sqrt(e[:, 0] * e[:, 0] + e[:, 1] * e[:, 1])
    nx, ny = -e[:, 1] / ln, e[:, 0] / ln
    c = (nx * v[:, 0] + ny * v[:, 1]) + d
    out = []
    for i in range(n):
        a, b = (i - 1) % n, i
        det = nx[a] * ny[b] - ny[a] * nx[b]
        out.append((float((c[a] * ny[b] - c[b] * ny[a]) / det), float((nx[a] * c[b] - nx[b] * c[a]) / det)))
    o = np.asarray(out)
    f = np.roll(o, -1, axis=0) - o
    if not np.all(e[:, 0] * f[:, 0] + e[:, 1] * f[:, 1] > 0.0):
        return None
    return out


def round_buffer_moments(poly: Sequence[Pt], r: float) -> Tuple[float, float, float]:
    """(A, A*cx, A*cy) of a convex polygon buffered by r with exact round joins (D2)."""
    poly = [tuple(map(float, p)) for p in poly]
    if shoelace(poly) < 0:
        poly = poly[::-1]
    n = len(poly)
    a0 = shoelace(poly)
    c0 = poly_centroid(poly)
    A, mx, my = a0, a0 * c0[0], a0 * c0[1]
    for k in range(n):
        x0, y0 = poly[k]
        x1, y1 = poly[(k + 1) % n]
        x2, y2 = poly[(k + 2) % n]
        ex, ey = x1 - x0, y1 - y0
        ln = math.hypot(ex, ey)
        ox, oy = ey / ln, -ex / ln
        ar = ln * r
        A += ar
        mx += ar * (0.5 * (x0 + x1) + 0.5 * r * ox)
        my += ar * (0.5 * (y0 + y1) + 0.5 * r * oy)
        fx, fy = x2 - x1, y2 - y1
        fl = math.hypot(fx, fy)
        o2x, o2y = fy / fl, -fx / fl
        phi = math.atan2(ox * o2y - oy * o2x, ox * o2x + oy * o2y)
        if phi > 0.0:
            sa = 0.5 * r * r * phi
            bx, by = ox + o2x, oy + o2y
            bl = math.hypot(bx, by)
            rad = 4.0 * r * math.sin(0.5 * phi) / (3.0 * phi)
            A += sa
            mx += sa * (x1 + rad * bx / bl)
            my += sa * (y1 + rad * by / bl)
    return A, mx, my


class QuadPolygon:
    """Simple polygon with optional holes (holes carry (area, mx, my) moments)."""

    def __init__(self, verts: Sequence[Pt], hole_moments: Sequence[Tuple[float, float, float]] = (),
                 inner: Optional["QuadPolygon"] = None):
        self._v = [tuple(map(float, p)) for p in verts]
        self._holes = list(hole_moments)
        self._inner = inner

    @property
    def is_empty(self) -> bool:
        return len(self._v) < 3

    @property
    def exterior(self):
        return _Ring(self._v)

    @property
    def interiors(self):
        return [] if self._inner is None else [_Ring(self._inner._v)]

    @property
    def bounds(self):
        v = np.asarray(self._v)
        return (float(v[:, 0].min()), float(v[:, 1].min()), float(v[:, 0].max()), float(v[:, 1].max()))

    @property
    def area(self) -> float:
        a = abs(shoelace(self._v))
        for h in self._holes:
            a -= h[0]
        if self._inner is not None:
            a -= self._inner.area
        return a

    @property
    def centroid(self):
        a = abs(shoelace(self._v))
        c = poly_centroid(self._v)
        mx, my = a * c[0], a * c[1]
        for h in self._holes:
            a -= h[0]
            mx -= h[1]
            my -= h[2]
        if self._inner is not None:
            ia = self._inner.area
            ic = self._inner.centroid
            a -= ia
            mx -= ia * ic.x
            my -= ia * ic.y
        return _Point(mx / a, my / a)

    def __repr__(self):
        return f"QuadPolygon({self._v}, area={self.area:.3f})"


def corner_angles_deg(verts: np.ndarray) -> np.ndarray:
    """Interior angles (degrees) of [F,4,2] quads — mlp3:165-192, vectorised."""
    prev = np.roll(verts, 1, axis=1) - verts
    nxt = np.roll(verts, -1, axis=1) - verts
    dot = prev[..., 0] * nxt[..., 0] + prev[..., 1] * nxt[..., 1]
    n1 = np.sqrt(prev[..., 0] ** 2 + prev[..., 1] ** 2)
    n2 = np.sqrt(nxt[..., 0] ** 2 + nxt[..., 1] ** 2)
    return np.degrees(np.arccos(np.clip(dot / (n1 * n2), -1.0, 1.0)))


def gap_gate(R: np.ndarray, W: float) -> np.ndarray:
    """``gap.area > 0.1`` (mlp3:1070, :1551) per candidate radius.

    gap = 2R x 2R square minus the W/2 round buffer of a 30-point quarter arc.  The whole
    buffer area, (pi/2)·R·W + pi·(W/2)², bounds the covered part from above, so the gate is
    decided analytically whenever that lower bound of the gap exceeds 0.1 m² (every sane (R, W));
    otherwise the area is evaluated numerically under D2."""
    R = np.asarray(R, dtype=np.float64)
    r = W / 2
    lower = 4 * R * R - (0.5 * math.pi * R * W + math.pi * r * r)
    gate = lower > 0.1
    for i in np.nonzero(~gate)[0]:
        gate[i] = _gap_area_numeric(float(R[i]), W) > 0.1
    return gate


def _gap_area_numeric(R: float, W: float, h: float = 0.01) -> float:
    a = np.linspace(0, np.pi / 2, 30)
    ax, ay = R * (1 - np.cos(a)), R * np.sin(a)       # corner 0 at the origin (mlp3:1127-1129)
    n = int(round(2 * R / h))
    xs = (np.arange(n) + 0.5) * h
    X, Y = np.meshgrid(xs, xs)
    cov = np.zeros(X.shape, dtype=bool)
    r2 = (W / 2) ** 2
    for k in range(len(a) - 1):
        dx, dy = ax[k + 1] - ax[k], ay[k + 1] - ay[k]
        wx, wy = X - ax[k], Y - ay[k]
        dd = dx * dx + dy * dy
        u = np.clip((wx * dx + wy * dy) / dd, 0.0, 1.0) if dd > 0 else 0.0
        qx, qy = wx - u * dx, wy - u * dy
        cov |= qx * qx + qy * qy < r2
    return float(np.count_nonzero(~cov)) * h * h
