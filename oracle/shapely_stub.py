"""Stand-in ``shapely`` / ``matplotlib`` modules so that the UNMODIFIED reference file
``/root/reference/multi_layer_planner_v3.py`` can be imported and executed in a container
that has neither (TEST INFRASTRUCTURE ONLY — see oracle/README.md).

Only the call sites the reference actually uses are implemented (SURVEY.md §8(c) lists
them), under the geometry decisions D1/D2 of oracle/geom.py:

* convex polygons only; ``buffer(-d)`` = mitred inset in input order (D1);
* ``buffer(+r)`` = exact round buffer, ``contains`` = ``dist < r`` strict (D2);
* areas of buffers/intersections/differences are NUMERIC (lattice sampling) — they only
  feed a gate (``gap.area > 0.1``), a print, and ``coverage_rate`` and are labelled
  "GEOS excluded" wherever they are stored.

Used by tests/golden/make_golden.py (fixture generation, run where /root/reference exists)
and by tests/test_oracle_vs_reference.py (skipped when /root/reference is absent).
"""
from __future__ import annotations

import sys
import types

import numpy as np

from . import geom


class _Coords(list):
    pass


class _PointLike:
    def __init__(self, x, y):
        self.x = float(x)
        self.y = float(y)
        self.coords = _Coords([(self.x, self.y)])


class Point(_PointLike):
    pass


class _Ring:
    def __init__(self, verts):
        self.coords = _Coords(list(verts) + [verts[0]])


class _Empty:
    is_empty = True
    area = 0.0
    bounds = ()

    def difference(self, other):
        return self


class Polygon:
    def __init__(self, coords):
        pts = [(float(p[0]), float(p[1])) for p in coords]
        if len(pts) > 1 and pts[0] == pts[-1]:
            pts = pts[:-1]
        self._v = pts

    # -- scalar properties -------------------------------------------------------------
    @property
    def is_empty(self):
        return len(self._v) < 3

    @property
    def area(self):
        return abs(geom.signed_area(self._v))

    @property
    def bounds(self):
        return geom.bounds(self._v)

    @property
    def centroid(self):
        return _PointLike(*geom.centroid(self._v))

    @property
    def exterior(self):
        return _Ring(self._v)

    # -- operations --------------------------------------------------------------------
    def buffer(self, d):
        if d < 0:
            ins = geom.inset_convex(self._v, -d)
            return _Empty() if ins is None else Polygon(ins)
        return RoundBuffer(self._segments(), d, fill=self)

    def _segments(self):
        v = self._v
        return [(v[i], v[(i + 1) % len(v)]) for i in range(len(v))]

    def _contains_xy(self, px, py, closed=True):
        """Convex CCW containment, numpy-broadcastable."""
        inside = np.ones(np.shape(px), dtype=bool)
        for (a, b) in self._segments():
            cr = (b[0] - a[0]) * (py - a[1]) - (b[1] - a[1]) * (px - a[0])
            inside &= (cr >= 0.0) if closed else (cr > 0.0)
        return inside

    def contains(self, pt):
        return bool(self._contains_xy(pt.x, pt.y, closed=False))

    def difference(self, other):
        if isinstance(other, Polygon):
            return PolygonWithHoles(self, [other])
        if isinstance(other, Union):
            return PolygonWithHoles(self, list(other.parts))
        if isinstance(other, RoundBuffer):
            return PolygonWithHoles(self, [other])
        raise TypeError(type(other))


class RoundBuffer:
    """Exact round buffer of a set of segments (and optionally a filled polygon)."""

    def __init__(self, segments, r, fill=None):
        self.segs = np.asarray([(a[0], a[1], b[0], b[1]) for a, b in segments], dtype=np.float64)
        self.r = float(r)
        self.fill = fill

    def _contains_xy(self, px, py):
        px = np.asarray(px, dtype=np.float64)
        py = np.asarray(py, dtype=np.float64)
        r2 = self.r * self.r
        inside = np.zeros(px.shape, dtype=bool)
        for ax, ay, bx, by in self.segs:
            inside |= geom.dist2_point_segment(px, py, ax, ay, bx, by) < r2
        if self.fill is not None:
            inside |= self.fill._contains_xy(px, py, closed=True)
        return inside

    def contains(self, pt):
        return bool(self._contains_xy(pt.x, pt.y))

    @property
    def bounds(self):
        xs = np.concatenate([self.segs[:, 0], self.segs[:, 2]])
        ys = np.concatenate([self.segs[:, 1], self.segs[:, 3]])
        return (xs.min() - self.r, ys.min() - self.r, xs.max() + self.r, ys.max() + self.r)

    def intersection(self, other):
        # cells inside the buffer are found segment by segment; only ``other`` is re-tested
        return _SampledRegion(other._contains_xy, other.bounds, segs=self.segs, r=self.r)


class Union:
    def __init__(self, parts):
        self.parts = list(parts)


class PolygonWithHoles:
    """outer polygon minus a list of convex polygons / round-buffered polygons."""

    def __init__(self, outer, holes):
        self.outer = outer
        self.holes = holes
        self.is_empty = False

    @property
    def bounds(self):
        return self.outer.bounds

    @property
    def exterior(self):
        return self.outer.exterior

    def _hole_moments(self, h):
        if isinstance(h, Polygon):
            a = h.area
            c = geom.centroid(h._v)
            return a, a * c[0], a * c[1]
        return geom.round_buffer_moments(h.fill._v, h.r)

    @property
    def area(self):
        a = self.outer.area
        if len(self.holes) == 1 and isinstance(self.holes[0], RoundBuffer) and self.holes[0].fill is None:
            # square minus buffered arc (corner gap, reference :1148): numeric
            return _SampledRegion(self._contains_xy, self.outer.bounds, h=0.02).area
        for h in self.holes:
            a -= self._hole_moments(h)[0]
        return a

    @property
    def centroid(self):
        a = self.outer.area
        c = geom.centroid(self.outer._v)
        mx, my = a * c[0], a * c[1]
        for h in self.holes:
            ha, hx, hy = self._hole_moments(h)
            a -= ha
            mx -= hx
            my -= hy
        return _PointLike(mx / a, my / a)

    def _contains_xy(self, px, py):
        inside = self.outer._contains_xy(px, py, closed=True)
        for h in self.holes:
            if isinstance(h, Polygon):
                inside &= ~h._contains_xy(px, py, closed=True)
            else:
                inside &= ~h._contains_xy(px, py)
        return inside


class _SampledRegion:
    """Area by cell-centre sampling (numeric; 'GEOS excluded')."""

    def __init__(self, pred, bnds, h=0.05, segs=None, r=None):
        self.pred = pred
        self.bnds = bnds
        self.h = h
        self.segs = segs
        self.r = r

    @property
    def area(self):
        x0, y0, x1, y1 = self.bnds
        h = self.h
        total = 0
        if self.segs is None:
            xs = x0 + (np.arange(int(round((x1 - x0) / h))) + 0.5) * h
            ys = y0 + (np.arange(int(round((y1 - y0) / h))) + 0.5) * h
            X, Y = np.meshgrid(xs, ys)
            return float(np.count_nonzero(self.pred(X, Y))) * h * h
        # sparse evaluation: only cells near some segment can be inside the buffer
        nx = int(round((x1 - x0) / h))
        ny = int(round((y1 - y0) / h))
        grid = np.zeros((ny, nx), dtype=bool)
        r = self.r
        for ax, ay, bx, by in self.segs:
            i0 = max(0, int(np.floor((min(ax, bx) - r - x0) / h)))
            i1 = min(nx, int(np.ceil((max(ax, bx) + r - x0) / h)) + 1)
            j0 = max(0, int(np.floor((min(ay, by) - r - y0) / h)))
            j1 = min(ny, int(np.ceil((max(ay, by) + r - y0) / h)) + 1)
            if i1 <= i0 or j1 <= j0:
                continue
            xs = x0 + (np.arange(i0, i1) + 0.5) * h
            ys = y0 + (np.arange(j0, j1) + 0.5) * h
            X, Y = np.meshgrid(xs, ys)
            grid[j0:j1, i0:i1] |= geom.dist2_point_segment(X, Y, ax, ay, bx, by) < r * r
        jj, ii = np.nonzero(grid)
        if len(jj) == 0:
            return 0.0
        X = x0 + (ii + 0.5) * h
        Y = y0 + (jj + 0.5) * h
        total = np.count_nonzero(self.pred(X, Y))
        return float(total) * h * h


class LineString:
    def __init__(self, coords):
        self._p = np.asarray(coords, dtype=np.float64).reshape(-1, 2)

    def buffer(self, r):
        p = self._p
        segs = [((p[i, 0], p[i, 1]), (p[i + 1, 0], p[i + 1, 1])) for i in range(len(p) - 1)]
        return RoundBuffer(segs, r)


def unary_union(parts):
    return Union(parts)


class _Dummy:
    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Dummy()

    def __getattr__(self, name):
        return _Dummy()

    def __iter__(self):
        return iter(())


def install():
    """Register the stub modules in ``sys.modules`` (idempotent; never overrides a real
    Shapely/matplotlib if one is importable)."""
    try:  # pragma: no cover - only on boxes that really have shapely
        import shapely.geometry  # noqa: F401
        have_shapely = True
    except Exception:
        have_shapely = False
    if not have_shapely:
        shp = types.ModuleType("shapely")
        g = types.ModuleType("shapely.geometry")
        o = types.ModuleType("shapely.ops")
        g.Polygon, g.LineString, g.Point = Polygon, LineString, Point
        o.unary_union = unary_union
        shp.geometry, shp.ops = g, o
        shp.__oracle_stub__ = True
        sys.modules["shapely"] = shp
        sys.modules["shapely.geometry"] = g
        sys.modules["shapely.ops"] = o
    try:  # pragma: no cover
        import matplotlib.pyplot  # noqa: F401
    except Exception:
        mpl = types.ModuleType("matplotlib")
        plt = types.ModuleType("matplotlib.pyplot")
        pat = types.ModuleType("matplotlib.patches")
        plt.__getattr__ = lambda name: _Dummy()  # type: ignore[attr-defined]
        pat.Polygon, pat.Rectangle = _Dummy, _Dummy
        mpl.pyplot, mpl.patches = plt, pat
        mpl.use = lambda *a, **k: None
        mpl.rcParams = {}
        sys.modules["matplotlib"] = mpl
        sys.modules["matplotlib.pyplot"] = plt
        sys.modules["matplotlib.patches"] = pat
    return not have_shapely


def load_reference(path="/root/reference/multi_layer_planner_v3.py"):
    """Import the unmodified reference planner module through the stubs."""
    import importlib.util
    import os

    if not os.path.exists(path):
        return None
    install()
    spec = importlib.util.spec_from_file_location("_reference_mlp3", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod
