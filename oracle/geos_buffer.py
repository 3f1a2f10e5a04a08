"""GEOS-faithful round buffer of a polyline, restated (TEST INFRASTRUCTURE ONLY) — SURVEY.md §8(f) N3.

The reference buffers its paths with Shapely's ``LineString.buffer(W / 2)`` (mlp3:1144, :1363, :1472, :1488):
GEOS with the default 16 segments per quadrant, round joins, round caps.  Shapely/GEOS cannot be installed here, so
decision D2 replaced that polygon by the EXACT round buffer (distance < r).  This module restates the REGION GEOS
builds, from its published algorithm (geos::operation::buffer::OffsetSegmentGenerator), to quantify what D2/D5 give
away:

  * every segment contributes its rectangle of half-width r (offset segments are exact straight lines);
  * an interior vertex with turn angle theta gets, on the OUTSIDE of the bend, a fillet from the end of the incoming
    offset segment to the start of the outgoing one: ``n = int(theta / (pi / 32) + 0.5)`` chords of equal angle
    (``addDirectedFillet``; n < 1 gives a straight bevel) with vertices ON the circle — an inscribed fan, so the
    GEOS region is a SUBSET of the exact round buffer and misses the circular slivers between the chords, at most
    r (1 - cos(inc / 2)) deep (1.93 mm for r = 1.6 m at the caps' pi/32 chords; up to 2.5 mm at the one-chord
    bevels of the 15-point corner arcs, whose 6.43 degree bends round to n = 1);
  * both ends get a half-circle cap of 32 chords;
  * (GEOS first drops input vertices closer than 1 % of r to the chord of their neighbours on the concave side —
    1.6 cm; the sampled arcs of the planner bend by 5 cm per vertex at R = 8 m, so nothing is dropped for R < 25 m
    and the concave side is bounded by exact offset lines either way.  Not modelled.)

``contains(path, r, X, Y)`` is ``buffer.contains(Point)`` for that region (interior only, like D2's strict <).
``sliver_area_bound`` is the total area of all slivers of a path (an upper bound of area(D2) - area(GEOS): slivers
that lie inside another segment's cover do not count).  **Parity with real GEOS stays unpinned** — this is a
restatement of the documented construction, not an execution of GEOS.
"""
from __future__ import annotations

import numpy as np

QUAD_SEGS = 16
QUANTUM = np.pi / 2 / QUAD_SEGS


def _clean(path):
    p = np.asarray(path, dtype=np.float64).reshape(-1, 2)
    keep = np.ones(len(p), dtype=bool)
    keep[1:] = np.any(p[1:] != p[:-1], axis=1)          # GEOS removes repeated points
    return p[keep]


def fillets(path):
    """[(centre, start angle, signed sweep, n chords)] of the joins and the two caps of a polyline."""
    p = _clean(path)
    out = []
    if len(p) < 2:
        return out
    d = p[1:] - p[:-1]
    ang = np.arctan2(d[:, 1], d[:, 0])
    # start cap: around p[0], from the left normal to the right normal through the back (counter-clockwise half turn)
    out.append((p[0], ang[0] + np.pi / 2, np.pi, 2 * QUAD_SEGS))
    for i in range(1, len(p) - 1):
        turn = (ang[i] - ang[i - 1] + np.pi) % (2 * np.pi) - np.pi          # > 0: left turn
        if turn == 0.0:
            continue
        theta = abs(turn)
        n = max(int(theta / QUANTUM + 0.5), 1)
        side = -np.pi / 2 if turn > 0 else np.pi / 2                        # the outside of the bend
        out.append((p[i], ang[i - 1] + side, turn, n))
    out.append((p[-1], ang[-1] - np.pi / 2, np.pi, 2 * QUAD_SEGS))
    return out


def contains(path, r, X, Y):
    """Boolean array: lattice points (X, Y) strictly inside the GEOS-style buffer of ``path``."""
    p = _clean(path)
    X = np.asarray(X, dtype=np.float64)
    Y = np.asarray(Y, dtype=np.float64)
    inside = np.zeros(X.shape, dtype=bool)
    if len(p) == 1:
        return inside
    for a, b in zip(p[:-1], p[1:]):                                          # rectangles
        dx, dy = b - a
        L = np.hypot(dx, dy)
        ux, uy = dx / L, dy / L
        t = (X - a[0]) * ux + (Y - a[1]) * uy
        s = (X - a[0]) * -uy + (Y - a[1]) * ux
        inside |= (t >= -1e-9) & (t <= L + 1e-9) & (np.abs(s) < r)     # (1 nm overlap: vertices belong to a rectangle)
    for c, a0, sweep, n in fillets(path):                                    # inscribed fans
        qx, qy = X - c[0], Y - c[1]
        rho = np.hypot(qx, qy)
        sgn = 1.0 if sweep > 0 else -1.0
        t = ((np.arctan2(qy, qx) - a0) * sgn) % (2 * np.pi)                  # angle from the fillet's first vertex
        inc = abs(sweep) / n
        k = np.floor(t / inc)
        mid = (k + 0.5) * inc
        inside |= (t <= abs(sweep)) & (rho * np.cos(t - mid) < r * np.cos(inc / 2))
    return inside


def exact_contains(path, r, X, Y):
    """Decision D2: distance to the polyline < r (the round buffer GEOS approximates), same float arithmetic."""
    p = _clean(path)
    X = np.asarray(X, dtype=np.float64)
    Y = np.asarray(Y, dtype=np.float64)
    d2 = np.full(X.shape, np.inf)
    for a, b in zip(p[:-1], p[1:]):
        dx, dy = b - a
        t = np.clip(((X - a[0]) * dx + (Y - a[1]) * dy) / (dx * dx + dy * dy), 0.0, 1.0)
        d2 = np.minimum(d2, (X - a[0] - t * dx) ** 2 + (Y - a[1] - t * dy) ** 2)
    return d2 < r * r, np.sqrt(d2)


def sliver_area_bound(path, r):
    """Sum over all fillets of (sector - inscribed fan): >= area(exact round buffer) - area(GEOS buffer)."""
    tot = 0.0
    for _, _, sweep, n in fillets(path):
        inc = abs(sweep) / n
        tot += n * 0.5 * r * r * (inc - np.sin(inc))
    return tot


def max_sliver_depth(path, r):
    return max((r * (1.0 - np.cos(abs(sweep) / n / 2)) for _, _, sweep, n in fillets(path)), default=0.0)


def raster_grid(path, r, ox, oy, h, nx, ny):
    """(geos, exact) boolean [ny, nx] grids of the lattice points (ox + i h, oy + j h): every segment and fillet only
    visits the lattice points of its own bounding box (whole-field grids in seconds)."""
    p = _clean(path)
    geos = np.zeros((ny, nx), dtype=bool)
    exact = np.zeros((ny, nx), dtype=bool)

    def block(x0, x1, y0, y1):
        i0, i1 = max(int(np.floor((x0 - ox) / h)), 0), min(int(np.ceil((x1 - ox) / h)) + 1, nx)
        j0, j1 = max(int(np.floor((y0 - oy) / h)), 0), min(int(np.ceil((y1 - oy) / h)) + 1, ny)
        if i0 >= i1 or j0 >= j1:
            return None
        X, Y = np.meshgrid(ox + np.arange(i0, i1) * h, oy + np.arange(j0, j1) * h)
        return (slice(j0, j1), slice(i0, i1)), X, Y

    for a, b in zip(p[:-1], p[1:]):
        blk = block(min(a[0], b[0]) - r, max(a[0], b[0]) + r, min(a[1], b[1]) - r, max(a[1], b[1]) + r)
        if blk is None:
            continue
        sl, X, Y = blk
        seg = np.array([a, b])
        geos[sl] |= contains(seg, r, X, Y) & ~_caps_only(seg, r, X, Y)
        exact[sl] |= exact_contains(seg, r, X, Y)[0]
    for c, a0, sweep, n in fillets(path):
        blk = block(c[0] - r, c[0] + r, c[1] - r, c[1] + r)
        if blk is None:
            continue
        sl, X, Y = blk
        qx, qy = X - c[0], Y - c[1]
        rho = np.hypot(qx, qy)
        sgn = 1.0 if sweep > 0 else -1.0
        t = ((np.arctan2(qy, qx) - a0) * sgn) % (2 * np.pi)
        inc = abs(sweep) / n
        mid = (np.floor(t / inc) + 0.5) * inc
        geos[sl] |= (t <= abs(sweep)) & (rho * np.cos(t - mid) < r * np.cos(inc / 2))
    return geos, exact


def _caps_only(seg, r, X, Y):
    """Lattice points that `contains` attributes to the CAPS of a single segment (not to its rectangle)."""
    a, b = seg
    dx, dy = b - a
    L = np.hypot(dx, dy)
    t = (X - a[0]) * dx / L + (Y - a[1]) * dy / L
    return (t < -1e-9) | (t > L + 1e-9)
