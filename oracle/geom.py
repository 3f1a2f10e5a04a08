"""Convex-polygon geometry used by the CPU oracle (TEST INFRASTRUCTURE ONLY).

This file is part of ``oracle/``: it may be imported only by ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs.  The product package never imports it.

It states the geometry decisions D1/D2 of SURVEY.md §8(c) that stand in for the
Shapely/GEOS calls of the reference (GEOS is not installable here, so everything
that depends on GEOS ring order or buffer discretisation is "parity unpinned"):

* D1  ``Polygon(quad).buffer(-d)``  -> exact mitred inset of a convex CCW quad, vertices
      returned in input order starting at input vertex 0
      (reference call sites: multi_layer_planner_v3.py:595, :871, :965).
* D2  ``geom.buffer(+r).contains(Point)`` -> exact ``dist(p, geom) < r`` (strict)
      (reference call sites: multi_layer_planner_v3.py:605, :1144, :1363, :1472, :1488,
      :1480, :1497).

Operation order in every formula below is normative: the CUDA kernels evaluate the same
expressions in the same order in FP64 with FMA contraction disabled, so integer results
derived from them agree bit for bit.
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence, Tuple

import numpy as np

Pt = Tuple[float, float]


def signed_area(verts: Sequence[Pt]) -> float:
    """Shoelace area (positive for CCW)."""
    s = 0.0
    n = len(verts)
    for i in range(n):
        x0, y0 = verts[i]
        x1, y1 = verts[(i + 1) % n]
        s += x0 * y1 - x1 * y0
    return 0.5 * s


def centroid(verts: Sequence[Pt]) -> Pt:
    """Area centroid of a simple polygon (what ``Polygon.centroid`` returns)."""
    n = len(verts)
    a2 = 0.0
    cx = 0.0
    cy = 0.0
    # shift to the first vertex for conditioning (fields may sit km from the origin)
    ox, oy = verts[0]
    for i in range(n):
        x0 = verts[i][0] - ox
        y0 = verts[i][1] - oy
        x1 = verts[(i + 1) % n][0] - ox
        y1 = verts[(i + 1) % n][1] - oy
        cr = x0 * y1 - x1 * y0
        a2 += cr
        cx += (x0 + x1) * cr
        cy += (y0 + y1) * cr
    return (ox + cx / (3.0 * a2), oy + cy / (3.0 * a2))


def bounds(verts: Sequence[Pt]) -> Tuple[float, float, float, float]:
    xs = [v[0] for v in verts]
    ys = [v[1] for v in verts]
    return (min(xs), min(ys), max(xs), max(ys))


def inset_convex(verts: Sequence[Pt], d: float) -> Optional[List[Pt]]:
    """D1: mitred inset of a convex CCW polygon by distance ``d`` (input order).

    New vertex i is the intersection of the inward-shifted lines of edge i-1 and edge i.
    Returns ``None`` when the inset collapses (an edge reverses direction) — the caller
    also applies the reference's ``area < 1.0`` rule (multi_layer_planner_v3.py:597, :967).
    """
    n = len(verts)
    nx = [0.0] * n
    ny = [0.0] * n
    c = [0.0] * n
    for k in range(n):
        x0, y0 = verts[k]
        x1, y1 = verts[(k + 1) % n]
        ex = x1 - x0
        ey = y1 - y0
        ln = math.sqrt(ex * ex + ey * ey)
        nx[k] = -ey / ln
        ny[k] = ex / ln
        c[k] = (nx[k] * x0 + ny[k] * y0) + d
    out: List[Pt] = []
    for i in range(n):
        a = (i - 1) % n
        b = i
        det = nx[a] * ny[b] - ny[a] * nx[b]
        px = (c[a] * ny[b] - c[b] * ny[a]) / det
        py = (nx[a] * c[b] - nx[b] * c[a]) / det
        out.append((px, py))
    # collapse check: every inset edge must keep the direction of its parent edge
    for k in range(n):
        ex = verts[(k + 1) % n][0] - verts[k][0]
        ey = verts[(k + 1) % n][1] - verts[k][1]
        fx = out[(k + 1) % n][0] - out[k][0]
        fy = out[(k + 1) % n][1] - out[k][1]
        if not (ex * fx + ey * fy > 0.0):
            return None
    return out


def rotate_point(p: Pt, cos_a: float, sin_a: float, center: Pt) -> Pt:
    """multi_layer_planner_v3.py:265-284 with the trig values passed in."""
    x = p[0] - center[0]
    y = p[1] - center[1]
    x_new = x * cos_a - y * sin_a
    y_new = x * sin_a + y * cos_a
    return (x_new + center[0], y_new + center[1])


def dist2_point_segment(px, py, ax, ay, bx, by):
    """Squared distance point->segment, numpy-broadcastable (D2).  Order is normative."""
    dx = bx - ax
    dy = by - ay
    wx = px - ax
    wy = py - ay
    dd = dx * dx + dy * dy
    t = wx * dx + wy * dy
    with np.errstate(divide="ignore", invalid="ignore"):
        u = np.where(dd > 0.0, t / np.where(dd > 0.0, dd, 1.0), 0.0)
    u = np.minimum(np.maximum(u, 0.0), 1.0)
    qx = wx - u * dx
    qy = wy - u * dy
    return qx * qx + qy * qy


def point_in_polygon_crossing(px, py, poly: np.ndarray):
    """Even-odd crossing test for a simple polygon; points on the boundary are undefined
    (callers combine it with a distance test).  numpy-broadcastable over px, py."""
    px = np.asarray(px, dtype=np.float64)
    py = np.asarray(py, dtype=np.float64)
    inside = np.zeros(px.shape, dtype=bool)
    n = len(poly)
    for k in range(n):
        ax, ay = poly[k]
        bx, by = poly[(k + 1) % n]
        cond = (ay > py) != (by > py)
        with np.errstate(divide="ignore", invalid="ignore"):
            xi = ax + (py - ay) * (bx - ax) / (by - ay)
        inside ^= cond & (px < xi)
    return inside


def round_buffer_moments(poly: Sequence[Pt], r: float) -> Tuple[float, float, float]:
    """Area and first moments (A, A*cx, A*cy) of a convex CCW polygon buffered by r under D2
    (exact round joins).  Used for the centroid of work-area-minus-obstacles
    (multi_layer_planner_v3.py:601-609, consumed only at :690 and :710)."""
    n = len(poly)
    a0 = signed_area(poly)
    c0 = centroid(poly)
    A = a0
    mx = a0 * c0[0]
    my = a0 * c0[1]
    for k in range(n):
        x0, y0 = poly[k]
        x1, y1 = poly[(k + 1) % n]
        ex = x1 - x0
        ey = y1 - y0
        ln = math.hypot(ex, ey)
        ox = ey / ln  # outward normal of a CCW polygon
        oy = -ex / ln
        ar = ln * r
        A += ar
        mx += ar * (0.5 * (x0 + x1) + 0.5 * r * ox)
        my += ar * (0.5 * (y0 + y1) + 0.5 * r * oy)
        # round join at vertex k+1 between this edge and the next
        x2, y2 = poly[(k + 2) % n]
        fx = x2 - x1
        fy = y2 - y1
        fl = math.hypot(fx, fy)
        o2x = fy / fl
        o2y = -fx / fl
        phi = math.atan2(ox * o2y - oy * o2x, ox * o2x + oy * o2y)  # exterior angle
        if phi > 0.0:
            sa = 0.5 * r * r * phi
            bx = ox + o2x
            by = oy + o2y
            bl = math.hypot(bx, by)
            rad = 4.0 * r * math.sin(0.5 * phi) / (3.0 * phi)
            A += sa
            mx += sa * (x1 + rad * bx / bl)
            my += sa * (y1 + rad * by / bl)
    return A, mx, my
